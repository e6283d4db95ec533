"""Tensor-level wrappers over the C-ABI (lora_b200._C) and the autograd glue.

torch is used for device memory, streams and autograd bookkeeping only; all arithmetic of the
LoRA path happens in liblora_b200.so. There is deliberately no eager fallback in this module.
"""
from typing import Optional, Tuple

import torch

from . import _C
from ._C import check, dtype_code, ptr, stream_ptr

R_PAD = 16
LAUNCH_COUNT = 0  # kernels of liblora_b200.so enqueued through this process (bench.py reads it)


def _count(n: int = 1):
    global LAUNCH_COUNT
    LAUNCH_COUNT += n


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _C.LoraB200Error("lora_b200 kernels need CUDA tensors (no CPU path)")


def cast_rows_pad16(src: torch.Tensor, rs: int, cs: int, r: int, C: int, dtype) -> torch.Tensor:
    """fp32 factor -> zero-padded 16-bit [16, C] operand (see lb_cast_rows_pad16)."""
    _req_cuda(src)
    out = torch.empty((R_PAD, C), device=src.device, dtype=dtype)
    check(_C.lib.lb_cast_rows_pad16(ptr(src), rs, cs, ptr(out), r, C, dtype_code(dtype),
                                    stream_ptr()), "lb_cast_rows_pad16")
    _count()
    return out


def cast_weight(w: torch.Tensor, dtype, want_plain: bool, want_t: bool):
    """Frozen weight [R,C] -> (16-bit copy or None, 16-bit transpose [C,R] or None)."""
    _req_cuda(w)
    if not want_plain and not want_t:
        return None, None
    w = w.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    R, C = w.shape
    plain = torch.empty((R, C), device=w.device, dtype=dtype) if want_plain else None
    wt = torch.empty((C, R), device=w.device, dtype=dtype) if want_t else None
    check(_C.lib.lb_cast_weight(ptr(w), dtype_code(w.dtype), ptr(plain), ptr(wt), R, C,
                                dtype_code(dtype), stream_ptr()), "lb_cast_weight")
    _count()
    return plain, wt


def fused_linear(x2d: torch.Tensor, w16: torch.Tensor, bias: Optional[torch.Tensor],
                 down16: torch.Tensor, up: torch.Tensor, up_rs: int, up_cs: int,
                 diag: Optional[torch.Tensor], scale: float, r: int, out_dtype,
                 want_t: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Y = x2d.w16^T + bias + ((x2d.down16^T) * scale*diag) . up^T ; returns (Y, T or None)."""
    _req_cuda(x2d, w16, down16, up)
    M, K = x2d.shape
    N = w16.shape[0]
    assert w16.shape[1] == K and down16.shape == (R_PAD, K)
    assert x2d.is_contiguous() and w16.is_contiguous() and down16.is_contiguous()
    y = torch.empty((M, N), device=x2d.device, dtype=out_dtype)
    t = torch.empty((M, R_PAD), device=x2d.device, dtype=torch.float32) if want_t else None
    check(_C.lib.lb_lora_linear_fwd(ptr(x2d), ptr(w16), ptr(bias), ptr(down16), ptr(up),
                                    up_rs, up_cs, ptr(diag), float(scale), ptr(y), ptr(t),
                                    M, K, N, r, dtype_code(x2d.dtype), dtype_code(out_dtype),
                                    stream_ptr()), "lb_lora_linear_fwd")
    _count()
    return y, t


def wgrad(S: torch.Tensor, V: torch.Tensor, diag: Optional[torch.Tensor], scale: float,
          out: torch.Tensor, out_js: int, out_cs: int, r: int):
    """out[j*out_js + c*out_cs] += scale*diag[j] * sum_m V[m,j]*S[m,c]."""
    _req_cuda(S, V, out)
    M, C = S.shape
    assert S.is_contiguous() and V.shape == (M, R_PAD) and V.dtype == torch.float32
    assert out.dtype == torch.float32
    check(_C.lib.lb_lora_wgrad(ptr(S), ptr(V), ptr(diag), float(scale), ptr(out), out_js, out_cs,
                               M, C, r, dtype_code(S.dtype), stream_ptr()), "lb_lora_wgrad")
    _count()
