"""Site discovery and LoRA injection (the reference's L2 "inject/find" API).

Behavioural contract = /root/reference/lora_diffusion/lora.py:159-380. The traversal ORDER is
part of the file format (safetensors keys are "{model}:{site_index}:up|down"), so discovery is
specified here exactly as the reference performs it:

  for every module `anc` of model.modules() (pre-order) whose CLASS NAME is in the target set:
      for every (dotted_name, m) of anc.named_modules() (pre-order):
          if m is an instance of a searched class, and its direct parent is not itself a LoRA
          operator module: emit (parent, child_name, m)

The generator is lazy in the reference and the tree is mutated while it is being consumed; the
same laziness is kept (an injected site is not re-emitted because its inner nn.Linear now has a
LoRA operator as direct parent).
"""
from typing import Iterable, Iterator, List, Optional, Sequence, Set, Tuple, Type

import weakref

import torch
import torch.nn as nn

from .modules import LoraInjectedConv2d, LoraInjectedLinear

UNET_DEFAULT_TARGET_REPLACE = {"CrossAttention", "Attention", "GEGLU"}
UNET_EXTENDED_TARGET_REPLACE = {"ResnetBlock2D", "CrossAttention", "Attention", "GEGLU"}
TEXT_ENCODER_DEFAULT_TARGET_REPLACE = {"CLIPAttention"}
TEXT_ENCODER_EXTENDED_TARGET_REPLACE = {"CLIPAttention"}
DEFAULT_TARGET_REPLACE = UNET_DEFAULT_TARGET_REPLACE
EMBED_FLAG = "<embed>"

Site = Tuple[nn.Module, str, nn.Module]


def _find_children(model, search_class: List[Type[nn.Module]] = [nn.Linear]) -> Iterator[Site]:
    """Direct children of any module that are instances of `search_class` (lora.py:172-186)."""
    wanted = tuple(search_class)
    for parent in model.modules():
        for name, child in parent.named_children():
            if isinstance(child, wanted):
                yield parent, name, child


def _find_modules_v2(
    model,
    ancestor_class: Optional[Set[str]] = None,
    search_class: List[Type[nn.Module]] = [nn.Linear],
    exclude_children_of: Optional[List[Type[nn.Module]]] = [LoraInjectedLinear, LoraInjectedConv2d],
) -> Iterator[Site]:
    """lora.py:189-232 (see module docstring for the exact order)."""
    wanted = tuple(search_class)
    banned = tuple(exclude_children_of) if exclude_children_of else ()
    if ancestor_class is None:
        roots: Iterable[nn.Module] = list(model.modules())
    else:
        roots = (m for m in model.modules() if type(m).__name__ in ancestor_class)
    for anc in roots:
        for dotted, mod in anc.named_modules():
            if not isinstance(mod, wanted):
                continue
            parts = dotted.split(".")
            parent = anc
            for hop in parts[:-1]:
                parent = parent.get_submodule(hop)
            if banned and isinstance(parent, banned):
                continue
            yield parent, parts[-1], mod


_find_modules = _find_modules_v2


def _adopt_frozen(dst_holder: nn.Module, src: nn.Module):
    """Share (not copy) the frozen Parameters of the original layer (lora.py:290-292)."""
    dst_holder.weight = src.weight
    if src.bias is not None:
        dst_holder.bias = src.bias


def _wrap_linear(child: nn.Linear, **kw) -> LoraInjectedLinear:
    new = LoraInjectedLinear(child.in_features, child.out_features, child.bias is not None, **kw)
    _adopt_frozen(new.linear, child)
    return new


def _wrap_conv(child: nn.Conv2d, **kw) -> LoraInjectedConv2d:
    new = LoraInjectedConv2d(child.in_channels, child.out_channels, child.kernel_size,
                             child.stride, child.padding, child.dilation, child.groups,
                             child.bias is not None, **kw)
    _adopt_frozen(new.conv, child)
    return new


def _finish_site(parent, name, new, loras, params, names):
    parent._modules[name] = new
    site = parent._modules[name]
    site._lb.parent = weakref.ref(parent)
    params.append(site.lora_up.parameters())
    params.append(site.lora_down.parameters())
    if loras is not None:
        # The reference assigns the list entries as they are (lora.py:302-303), which only works
        # when the .pt holds nn.Parameter objects; files written by save_lora_weight hold plain
        # fp16 tensors and make that assignment raise TypeError. Accept both: plain tensors are
        # wrapped (and cast to the site's dtype, as monkeypatch_or_replace_lora does, lora.py:706-711).
        like = site.lora_up.weight
        up, down = loras.pop(0), loras.pop(0)
        if not isinstance(up, nn.Parameter):
            up = nn.Parameter(up.detach().to(like.device, like.dtype))
        if not isinstance(down, nn.Parameter):
            down = nn.Parameter(down.detach().to(like.device, like.dtype))
        site.lora_up.weight = up
        site.lora_down.weight = down
    site.lora_up.weight.requires_grad = True
    site.lora_down.weight.requires_grad = True
    names.append(name)


def inject_trainable_lora(
    model: nn.Module,
    target_replace_module: Set[str] = DEFAULT_TARGET_REPLACE,
    r: int = 4,
    loras=None,  # path to a .pt list [up0, down0, up1, ...]
    verbose: bool = False,
    dropout_p: float = 0.0,
    scale: float = 1.0,
):
    """Swap every nn.Linear under a target ancestor for a LoraInjectedLinear (lora.py:255-309).
    Returns ([param generators: up0, down0, up1, ...], [child names])."""
    params, names = [], []
    if loras is not None:
        loras = torch.load(loras)
    for parent, name, child in _find_modules(model, target_replace_module, search_class=[nn.Linear]):
        if verbose:
            print("LoRA Injection : injecting lora into ", name)
            print("LoRA Injection : weight shape", child.weight.shape)
        new = _wrap_linear(child, r=r, dropout_p=dropout_p, scale=scale)
        new.to(child.weight.device).to(child.weight.dtype)
        _finish_site(parent, name, new, loras, params, names)
    return params, names


def inject_trainable_lora_extended(
    model: nn.Module,
    target_replace_module: Set[str] = UNET_EXTENDED_TARGET_REPLACE,
    r: int = 4,
    loras=None,
):
    """Linear AND Conv2d sites (lora.py:312-380). No dropout/scale arguments in the reference:
    the operator defaults (dropout_p = 0.1, scale = 1.0) apply."""
    params, names = [], []
    if loras is not None:
        loras = torch.load(loras)
    for parent, name, child in _find_modules(model, target_replace_module,
                                             search_class=[nn.Linear, nn.Conv2d]):
        if type(child) is nn.Linear:
            new = _wrap_linear(child, r=r)
        elif type(child) is nn.Conv2d:
            new = _wrap_conv(child, r=r)
        else:
            # subclass of Linear/Conv2d: the reference falls through with a stale `_tmp`; we skip
            continue
        new.to(child.weight.device).to(child.weight.dtype)
        if child.bias is not None:
            new.to(child.bias.device).to(child.bias.dtype)
        _finish_site(parent, name, new, loras, params, names)
    return params, names


def extract_lora_ups_down(model, target_replace_module=DEFAULT_TARGET_REPLACE):
    """[(lora_up module, lora_down module)] in site order (lora.py:383-397)."""
    pairs = [(m.lora_up, m.lora_down) for _, _, m in _find_modules(
        model, target_replace_module, search_class=[LoraInjectedLinear, LoraInjectedConv2d])]
    if not pairs:
        raise ValueError("No lora injected.")
    return pairs


def extract_lora_as_tensor(model, target_replace_module=DEFAULT_TARGET_REPLACE, as_fp16=True):
    """[(up * scale, down)] tensors, fp16 by default (lora.py:400-421)."""
    out = []
    for _, _, m in _find_modules(model, target_replace_module,
                                 search_class=[LoraInjectedLinear, LoraInjectedConv2d]):
        up, down = m.realize_as_lora()
        if as_fp16:
            up, down = up.to(torch.float16), down.to(torch.float16)
        out.append((up, down))
    if not out:
        raise ValueError("No lora injected.")
    return out
