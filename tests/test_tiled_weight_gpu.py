"""LB_W_TILED: the frozen weight stored as contiguous 64 x 64 blocks (lb_tile_weight) gives the same
results as the row-major operand in every tile schedule, and the tiler matches a torch restatement."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _torch_tiled(w, dtype):
    N, K = w.shape
    n64, nkb = (N + 63) // 64, (K + 63) // 64
    pad = torch.zeros(n64 * 64, nkb * 64, device=w.device, dtype=torch.float32)
    pad[:N, :K] = w.float()
    return pad.view(n64, 64, nkb, 64).permute(0, 2, 1, 3).contiguous().to(dtype).reshape(-1)


@pytest.mark.parametrize("src_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,K", [(320, 768), (200, 72), (64, 64), (1000, 136)])
def test_tile_weight_matches_torch(N, K, src_dtype):
    from lora_b200 import ops
    torch.manual_seed(N + K)
    w = torch.randn(N, K, device=DEV).to(src_dtype)
    tw = ops.tile_weight(w, torch.bfloat16)
    assert tw.shape == (N, K) and torch.equal(tw.buf, _torch_tiled(w, torch.bfloat16))
    twt = ops.tile_weight(w, torch.bfloat16, transpose=True)
    assert twt.shape == (K, N) and torch.equal(twt.buf, _torch_tiled(w.t().contiguous(), torch.bfloat16))


@pytest.mark.parametrize("mode", [0, 1 + 4, 1 + 8, 2 + 4, 2 + 8, 1 + 12])
@pytest.mark.parametrize("M,K,N,r", [(77, 768, 320, 4), (256, 1280, 2304, 4), (300, 328, 200, 8), (64, 10240, 1280, 4),
                                     (2048, 320, 1280, 16)])
def test_fused_linear_with_tiled_weight_equals_row_major(M, K, N, r, mode):
    from lora_b200 import _C, ops
    torch.manual_seed(M + K + N)
    dt = torch.bfloat16
    x = torch.randn(M, K, device=DEV, dtype=dt)
    w = (torch.randn(N, K, device=DEV) * 0.03).to(dt)
    a = torch.randn(r, K, device=DEV)
    b = torch.randn(N, r, device=DEV) * 0.05
    bias = torch.randn(N, device=DEV)
    d16 = ops.cast_rows_pad16(a, K, 1, r, K, dt)
    tw = ops.tile_weight(w, dt)
    try:
        assert _C.lib.lb_debug_set_linear_mode(mode) == 0
        y0, t0 = ops.fused_linear(x, w, bias, d16, b, r, 1, None, 0.7, r, dt, True)
        y1, t1 = ops.fused_linear(x, tw, bias, d16, b, r, 1, None, 0.7, r, dt, True)
    finally:
        _C.lib.lb_debug_set_linear_mode(0)
    torch.cuda.synchronize()
    assert torch.equal(t0, t1) or float((t0 - t1).norm() / t0.norm()) < 3e-6    # split-K plans: summation order
    assert float((y0.float() - y1.float()).norm() / y0.float().norm()) < 2e-3
