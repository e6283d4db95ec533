"""ORACLE (test infrastructure only -- never imported by the product path).

PyTorch-eager restatement of the reference's LoRA operator modules and of its site traversal,
used (a) as the CPU baseline timed beside the CUDA kernels (bench.py --impl reference /
cpu_baseline) and (b) as the module-level parity target on the GPU box, where /root/reference
does not exist. Checked against the real /root/reference/lora_diffusion/lora.py in
tests/test_vs_reference_live.py (runs wherever the reference tree is mounted) and through
the golden vectors it generated (tests/golden/, scripts/make_golden.py).

Written functionally (F.linear / F.conv2d on explicit Parameters) rather than as the
reference's nested nn.Linear children; the arithmetic and its ORDER are the reference's:
    lora_diffusion/lora.py:53-58   y = F.linear(x, W, b) + dropout(up(selector(down(x)))) * scale
    lora_diffusion/lora.py:130-135 same with conv2d (down: base geometry, up: 1x1)
"""
from typing import Iterator, List, Optional, Set, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class RefLoraSite(nn.Module):
    """One LoRA site around a frozen nn.Linear or nn.Conv2d (weights shared, not copied)."""

    def __init__(self, base: nn.Module, r: int = 4, dropout_p: float = 0.1, scale: float = 1.0):
        super().__init__()
        self.is_conv = isinstance(base, nn.Conv2d)
        lo = min(base.in_channels, base.out_channels) if self.is_conv else min(base.in_features, base.out_features)
        if r > lo:
            raise ValueError(f"LoRA rank {r} must be less or equal than {lo}")
        self.r = r
        self.scale = scale
        self.p = dropout_p
        self.weight = base.weight          # frozen, shared (lora.py:290-292)
        self.bias = base.bias
        if self.is_conv:
            self.geom = dict(stride=base.stride, padding=base.padding, dilation=base.dilation,
                             groups=base.groups)
            kh, kw = base.kernel_size
            down_shape = (r, base.in_channels // base.groups, kh, kw)
            up_shape = (base.out_channels, r, 1, 1)
        else:
            down_shape = (r, base.in_features)
            up_shape = (base.out_features, r)
        w = base.weight
        self.down = nn.Parameter(torch.empty(down_shape, device=w.device, dtype=w.dtype))
        self.up = nn.Parameter(torch.zeros(up_shape, device=w.device, dtype=w.dtype))
        nn.init.normal_(self.down, std=1 / r)      # lora.py:50-51 / 127-128
        self.diag: Optional[torch.Tensor] = None   # selector (lora.py:63-70)

    def forward(self, x):
        if self.is_conv:
            base = F.conv2d(x, self.weight, self.bias, **self.geom)
            t = F.conv2d(x, self.down, None, **self.geom)
            if self.diag is not None:
                t = t * self.diag.view(1, -1, 1, 1).to(t.dtype)
            u = F.conv2d(t, self.up)
        else:
            base = F.linear(x, self.weight, self.bias)
            t = F.linear(x, self.down)
            if self.diag is not None:
                t = t * self.diag.to(t.dtype)
            u = F.linear(t, self.up)
        u = F.dropout(u, self.p, self.training)
        return base + u * self.scale


def ref_find_sites(model: nn.Module, ancestor_names: Set[str],
                   kinds: Tuple[type, ...]) -> Iterator[Tuple[nn.Module, str, nn.Module]]:
    """lora.py:189-232: ancestors in model.modules() order (matched by class-name string), then
    ancestor.named_modules() order; children whose direct parent is a LoRA site are skipped."""
    for anc in (m for m in model.modules() if m.__class__.__name__ in ancestor_names):
        for dotted, mod in anc.named_modules():
            if not isinstance(mod, kinds):
                continue
            *path, leaf = dotted.split(".")
            parent = anc
            for hop in path:
                parent = parent.get_submodule(hop)
            if isinstance(parent, RefLoraSite):
                continue
            yield parent, leaf, mod


def ref_inject(model: nn.Module, targets: Set[str], r: int = 4, dropout_p: float = 0.0,
               scale: float = 1.0, extended: bool = False) -> List[RefLoraSite]:
    """inject_trainable_lora (lora.py:255-309; dropout 0.0) or, with extended=True,
    inject_trainable_lora_extended (lora.py:312-380; Linear+Conv2d, class-default dropout 0.1)."""
    kinds = (nn.Linear, nn.Conv2d) if extended else (nn.Linear,)
    if extended:
        dropout_p = 0.1
    sites = []
    for parent, name, child in ref_find_sites(model, targets, kinds):
        site = RefLoraSite(child, r=r, dropout_p=dropout_p, scale=scale)
        parent._modules[name] = site
        sites.append(site)
    return sites
