"""EXPERIMENTAL cluster split-K schedule of the fused LoRA linear kernel (csrc/fused_splitk.cuh).

Opt-in: runs only with LB_EXPERIMENTAL=1 in the environment. The kernel was written after round
1's GPU budget was spent (compiled and SASS-inspected, never executed); a protocol bug in it would
`__trap` and poison the CUDA context of the whole pytest process, so it must not run by default.
    LB_EXPERIMENTAL=1 python -m pytest tests/test_zz_splitk_experimental_gpu.py -x -q
Same acceptance as tests/test_kernels_gpu.py::test_every_tile_schedule_gives_the_same_result."""
import os

import pytest
import torch

from oracle import lora_ops as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("LB_EXPERIMENTAL") != "1",
                                 reason="experimental kernel: set LB_EXPERIMENTAL=1 to run")]

SHAPES = [(77, 768, 768, 4), (77, 768, 320, 4), (256, 1280, 1280, 4), (64, 1280, 1280, 8),
          (1024, 640, 640, 4), (256, 10240, 1280, 4), (300, 200, 136, 16), (128, 64, 64, 1)]


@pytest.mark.parametrize("split", [2, 3, 4])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,K,N,r", SHAPES)
def test_cluster_splitk_matches_default_schedule(M, K, N, r, split, out_dtype):
    from lora_b200 import _C
    from test_kernels_gpu import make_case, rel_err, run_fused     # tests/ is on sys.path under pytest
    x, W, A, B, b, d = make_case(M, K, N, r, torch.bfloat16, seed=M + K + N + r, diag=True)
    mode = 3 + 4 * 1 + 16 * (split - 1)           # split-K schedule, BLOCK_N 64, `split` CTAs per tile
    try:
        assert _C.lib.lb_debug_set_linear_mode(mode) == 0
        y, t, _ = run_fused(x, W, A, B, b, d, 0.9, out_dtype)
    finally:
        _C.lib.lb_debug_set_linear_mode(0)
    y0, t0, _ = run_fused(x, W, A, B, b, d, 0.9, out_dtype)
    assert rel_err(t, t0) < 3e-6
    assert rel_err(y, y0) < (3e-4 if out_dtype == torch.float32 else 1e-2)
    if out_dtype == torch.float32:
        ref = O.lora_linear_forward(x, W, b, A, B, 0.9, diag=d)
        branch = (ref - O.lora_linear_forward(x, W, b, A, torch.zeros_like(B), 0.0)).norm()
        assert float((y.double().cpu() - ref).norm()) <= 2.0 ** -7 * float(branch) + 1e-5 * float(ref.norm())


def test_cluster_splitk_wide_tiles():
    from lora_b200 import _C
    from test_kernels_gpu import make_case, rel_err, run_fused     # tests/ is on sys.path under pytest
    x, W, A, B, b, d = make_case(256, 10240, 1280, 4, torch.bfloat16, seed=5, diag=False)
    try:
        assert _C.lib.lb_debug_set_linear_mode(3 + 4 * 2 + 16 * 1) == 0    # BLOCK_N 128, 2 CTAs per tile
        y, t, _ = run_fused(x, W, A, B, b, d, 1.0, torch.bfloat16)
    finally:
        _C.lib.lb_debug_set_linear_mode(0)
    y0, t0, _ = run_fused(x, W, A, B, b, d, 1.0, torch.bfloat16)
    assert rel_err(t, t0) < 3e-6 and rel_err(y, y0) < 1e-2
