"""SVD distillation of a fine-tune into LoRA factors, batched on the GPU.

Contract = /root/reference/lora_diffusion/cli_svd.py:24-142 (`overwrite_base`, `svd_distill`):
per LoRA site, dW = (W_tuned - W_base) (conv: flattened to [Cout, Cin*kh*kw]), keep the top `rank`
singular triplets, up = U_r diag(S_r), down = Vh_r, clamp both symmetrically at the
`clamp_quantile` quantile of their joint values, write them into lora_up / lora_down.

The reference loops over sites calling a full `torch.linalg.svd`. Here all same-shape sites are
processed together by the primitives in csrc/svd.cu (randomized range finder, 32 probes,
q power iterations re-orthonormalised through a 32x32 Jacobi eigen-solve): the only passes over the weights are
2(q+1) streaming reads of (W_tuned, W_base).

Singular vectors are unique up to a per-component sign, and the reference's clamp threshold is a
quantile over SIGNED values, so elementwise equality with LAPACK is not defined;
parity is stated on singular values, on the rank-r product before clamping and on the clamp rule
(tests/test_svd_gpu.py).
"""
import ctypes
from collections import defaultdict
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _C, ops
from ._C import check, dtype_code, ptr, stream_ptr

L = 32


def _ptr_array(tensors: Sequence[torch.Tensor], device) -> torch.Tensor:
    return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device=device)


def _orth2(Y: torch.Tensor, G: torch.Tensor, V: torch.Tensor, sig: torch.Tensor, rows: int, batch: int,
           sweeps: int = 6):
    """Orthonormalise the 32 columns of Y in place, twice (like CholeskyQR2, but rank-revealing):
    G = Y^T Y = V diag(s^2) V^T (Jacobi)  ->  Y <- Y V diag(1/s), numerically null directions
    (s < 1e-6 s_max) become zero columns. A Cholesky factor would break down exactly there, and
    exactly-low-rank deltas are a real input (a fine-tune that IS a merged LoRA)."""
    lib, st = _C.lib, stream_ptr()
    for _ in range(2):
        check(lib.lb_svd_gram(ptr(Y), ptr(G), rows, batch, st), "lb_svd_gram")
        check(lib.lb_svd_jacobi(ptr(G), ptr(V), ptr(sig), batch, sweeps, st), "lb_svd_jacobi")
        check(lib.lb_svd_apply(ptr(Y), ptr(V), ptr(sig), 2, ptr(Y), rows, L, L, 0, rows * L, batch, st),
              "lb_svd_apply")
        ops._count(3)


def svd_lowrank_batched(W_tuned: Sequence[torch.Tensor], W_base: Sequence[torch.Tensor], rank: int,
                        power_iters: int = 2, seed: int = 0, jacobi_sweeps: int = 10
                        ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Same-shape 2-D weight pairs -> (up [b,N,rank] = U_r diag(S_r), down [b,rank,K] = Vh_r,
    sigma [b,32] descending). Weights may be fp32 / bf16 / fp16 (all the same dtype)."""
    if not 1 <= rank <= 16:
        raise _C.LoraB200Error("svd_lowrank_batched: rank must be in [1,16]")
    batch = len(W_tuned)
    w0 = W_tuned[0]
    if not w0.is_cuda:
        raise _C.LoraB200Error("svd_lowrank_batched needs CUDA tensors (no CPU path)")
    N, K = w0.shape
    dev = w0.device
    keep = [t.detach().contiguous() for t in list(W_tuned) + list(W_base)]
    for t in keep:
        assert t.shape == (N, K) and t.dtype == w0.dtype
    pt, pb = _ptr_array(keep[:batch], dev), _ptr_array(keep[batch:], dev)
    wdt = dtype_code(w0.dtype)
    lib, st = _C.lib, stream_ptr()
    f32 = dict(device=dev, dtype=torch.float32)
    Z = torch.empty((batch, K, L), **f32)
    Y = torch.empty((batch, N, L), **f32)
    G = torch.empty((batch, L, L), **f32)
    V = torch.empty((batch, L, L), **f32)
    sigma = torch.empty((batch, L), **f32)

    check(lib.lb_svd_randn(ptr(Z), Z.numel(), ctypes.c_ulonglong(seed * 2654435761 + 12345), st), "lb_svd_randn")
    check(lib.lb_svd_mul(ptr(pt), ptr(pb), wdt, ptr(Z), ptr(Y), N, K, batch, 0, st), "lb_svd_mul")
    ops._count(2)
    _orth2(Y, G, V, sigma, N, batch)
    for _ in range(power_iters):
        check(lib.lb_svd_mul(ptr(pt), ptr(pb), wdt, ptr(Y), ptr(Z), N, K, batch, 1, st), "lb_svd_mul")
        _orth2(Z, G, V, sigma, K, batch)
        check(lib.lb_svd_mul(ptr(pt), ptr(pb), wdt, ptr(Z), ptr(Y), N, K, batch, 0, st), "lb_svd_mul")
        _orth2(Y, G, V, sigma, N, batch)
        ops._count(2)
    # Q = Y (orthonormal); Z = dW^T Q = B^T ;  B B^T = Z^T Z = Uh diag(s^2) Uh^T
    check(lib.lb_svd_mul(ptr(pt), ptr(pb), wdt, ptr(Y), ptr(Z), N, K, batch, 1, st), "lb_svd_mul")
    check(lib.lb_svd_gram(ptr(Z), ptr(G), K, batch, st), "lb_svd_gram")
    check(lib.lb_svd_jacobi(ptr(G), ptr(V), ptr(sigma), batch, jacobi_sweeps, st), "lb_svd_jacobi")
    up = torch.empty((batch, N, rank), **f32)
    down = torch.empty((batch, rank, K), **f32)
    # up = (Q Uh)[:, :r] * s ;  down = ((Z Uh)[:, :r] / s)^T
    check(lib.lb_svd_apply(ptr(Y), ptr(V), ptr(sigma), 1, ptr(up), N, rank, rank, 0, N * rank, batch, st),
          "lb_svd_apply")
    check(lib.lb_svd_apply(ptr(Z), ptr(V), ptr(sigma), 2, ptr(down), K, rank, K, 1, rank * K, batch, st),
          "lb_svd_apply")
    ops._count(5)
    return up, down, sigma


def _clamp_pair_(up: torch.Tensor, down: torch.Tensor, q: float):
    """cli_svd.py:42-47: hi = quantile(cat(U, Vh), q); clamp both to [-hi, hi] (in place)."""
    hi = torch.quantile(torch.cat([up.flatten(), down.flatten()]), q)
    up.clamp_(-hi, hi)
    down.clamp_(-hi, hi)
    return hi


def _iter_lora(model):
    for m in model.modules():
        if type(m).__name__ in ("LoraInjectedLinear", "LoraInjectedConv2d"):
            yield m


def overwrite_base(base_model: nn.Module, tuned_model: nn.Module, rank: int, clamp_quantile: float,
                   power_iters: int = 2, shard: Optional[Tuple[int, int]] = None):
    """cli_svd.py:24-92 on two LoRA-injected models with identical site structure: the factors of
    `base_model`'s sites are overwritten with the rank-`rank` distillation of (tuned - base).

    shard=(rank, world) (SURVEY.md 8e, config C5 on several GPUs): the weight deltas are
    independent, so this process distils only sites i = rank (mod world) in traversal order -- no
    collective on the compute path -- and the factors are then exchanged with ONE SUM all-reduce
    over a zero-initialised flat fp32 buffer in which every rank has filled only its own slots
    (exact: each slot receives one value plus zeros). Every rank ends with all sites overwritten."""
    sites = list(zip(_iter_lora(base_model), _iter_lora(tuned_model)))
    me, world = shard if shard is not None else (0, 1)
    groups = defaultdict(list)
    for i, (sb, st_) in enumerate(sites):
        if i % world != me:
            continue
        hb = sb.linear if hasattr(sb, "linear") else sb.conv
        ht = st_.linear if hasattr(st_, "linear") else st_.conv
        wb, wt = hb.weight.data, ht.weight.data
        groups[(tuple(wb.shape), wb.dtype)].append((sb, wb.flatten(start_dim=1), wt.flatten(start_dim=1)))
    for (shape, _), items in groups.items():
        ups, downs, _ = svd_lowrank_batched([x[2] for x in items], [x[1] for x in items], rank,
                                            power_iters=power_iters)
        # cli_svd.py:42-47 for the whole group at once: hi_b = quantile(cat(U_b, Vh_b), q); clamp to [-hi_b, hi_b]
        b = ups.shape[0]
        hi = torch.quantile(torch.cat([ups.reshape(b, -1), downs.reshape(b, -1)], dim=1), clamp_quantile, dim=1)
        ups = torch.minimum(torch.maximum(ups, -hi.view(b, 1, 1)), hi.view(b, 1, 1))
        downs = torch.minimum(torch.maximum(downs, -hi.view(b, 1, 1)), hi.view(b, 1, 1))
        for i, (site, _, _) in enumerate(items):
            u, d = ups[i], downs[i]
            dev, dt = site.lora_up.weight.device, site.lora_up.weight.dtype
            assert site.lora_up.weight.flatten(1).shape == u.shape
            assert site.lora_down.weight.flatten(1).shape == d.shape
            site.lora_up.weight.data = u.reshape(site.lora_up.weight.shape).to(device=dev, dtype=dt)
            site.lora_down.weight.data = d.reshape(site.lora_down.weight.shape).to(device=dev, dtype=dt)
    if world > 1:
        _exchange_factors([sb for sb, _ in sites], me, world)


def _exchange_factors(sites, me: int, world: int):
    """Every rank holds valid factors for sites i = me (mod world); after this every rank holds all."""
    from .dist import allreduce_sum_
    tensors = [t for s in sites for t in (s.lora_up.weight, s.lora_down.weight)]
    owners = [i % world for i in range(len(sites)) for _ in range(2)]
    flat = torch.zeros(sum(t.numel() for t in tensors), dtype=torch.float32, device=tensors[0].device)
    off = 0
    for t, owner in zip(tensors, owners):
        if owner == me:
            flat[off:off + t.numel()] = t.data.flatten().float()
        off += t.numel()
    allreduce_sum_(flat)
    off = 0
    for t in tensors:
        t.data = flat[off:off + t.numel()].reshape(t.shape).to(t.dtype)
        off += t.numel()


def svd_distill(unet_base: nn.Module, unet_tuned: nn.Module, text_base: nn.Module, text_tuned: nn.Module,
                rank: int = 4, clamp_quantile: float = 0.99, save_path: str = "svd_distill.safetensors",
                shard: Optional[Tuple[int, int]] = None):
    """cli_svd.py:95-142 with module arguments instead of hub ids (diffusers pipelines cannot be
    loaded offline): inject (extended for the UNet, CLIPAttention for the text encoder), distill,
    save with save_all (which, like the reference, writes only the default-target UNet sites)."""
    from .inject import inject_trainable_lora, inject_trainable_lora_extended
    from .persist import save_all
    inject_trainable_lora_extended(unet_base, r=rank)
    inject_trainable_lora_extended(unet_tuned, r=rank)
    overwrite_base(unet_base, unet_tuned, rank=rank, clamp_quantile=clamp_quantile, shard=shard)
    inject_trainable_lora(text_base, r=rank, target_replace_module={"CLIPAttention"})
    inject_trainable_lora(text_tuned, r=rank, target_replace_module={"CLIPAttention"})
    overwrite_base(text_base, text_tuned, rank=rank, clamp_quantile=clamp_quantile, shard=shard)
    if shard is not None and shard[0] != 0:
        return                              # sharded run: every rank holds all factors, rank 0 writes
    save_all(unet=unet_base, text_encoder=text_base, placeholder_token_ids=None, placeholder_tokens=None,
             save_path=save_path, save_lora=True, save_ti=False)
