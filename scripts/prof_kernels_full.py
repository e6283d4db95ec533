"""One launch of every hand-written kernel of the path inside a cudaProfilerStart/Stop range, for
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/r2_kernels python scripts/prof_kernels_full.py
(one GPU; ~20 kernels x ~40 replays). Summarised here with scripts/summarize_ncu_full.py into
profiles/r2_ncu_per_kernel.md. Shapes are real SD1.5 site shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import lora_b200 as L
from lora_b200 import ops
from lora_b200.arena import LoraArena

dev, dt = "cuda", torch.bfloat16
torch.manual_seed(0)


def lin_case(M, K, N, r=4):
    x = torch.randn(M, K, device=dev, dtype=dt)
    w = torch.randn(N, K, device=dev, dtype=dt) * 0.02
    a = torch.randn(r, K, device=dev)
    b = torch.randn(N, r, device=dev) * 0.01
    d16 = ops.cast_rows_pad16(a, K, 1, r, K, dt)
    bias = torch.zeros(N, device=dev)
    return x, w, bias, d16, b, r


cases = []
# fused linear: persistent (GEGLU fwd), one tile per CTA, split-K (GEGLU dX), grouped q/k/v
# + the cluster split-K kernel (256x1280->1280) and the 192-wide single-wave tile (256x1280->10240)
for shape in ((4096, 320, 2560), (1024, 640, 640), (256, 10240, 1280), (256, 1280, 1280), (256, 1280, 10240)):
    x, w, bias, d16, b, r = lin_case(*shape)
    cases.append(lambda x=x, w=w, bias=bias, d16=d16, b=b, r=r: ops.fused_linear(x, w, bias, d16, b, r, 1, None, 1.0, r, dt, True))
xq = torch.randn(1024, 640, device=dev, dtype=dt)
probs = []
for _ in range(3):
    _, w, bias, d16, b, r = lin_case(1024, 640, 640)
    probs.append((xq, w, None, d16, b, r, 1, None, 1.0, r))
cases.append(lambda: ops.fused_linear_grouped(probs, dt, True))
# dropout forward (mask in the drain)
x, w, bias, d16, b, r = lin_case(4096, 320, 320)
seed = torch.zeros(1, device=dev, dtype=torch.int64)
cases.append(lambda: ops.fused_linear(x, w, bias, d16, b, r, 1, None, 1.0, r, dt, True, drop_p=0.1, seed=seed))
# conv: forward with dropout, input gradient (9 T groups), 8x8 split-K forward
for (cin, cout, H) in ((320, 320, 64), (1280, 1280, 8)):
    xc = torch.randn(1, cin, H, H, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
    gyc = torch.randn(1, cout, H, H, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
    wc = torch.randn(cout, cin, 3, 3, device=dev) * 0.02
    wf, wb = ops.cast_conv_weight(wc, dt, True, True)
    A = torch.randn(8, cin, 3, 3, device=dev)
    B = torch.randn(cout, 8, device=dev) * 0.01
    cd16 = ops.conv_down16(A, dt, {})
    bt16 = ops.cast_rows_pad16(B, 1, 8, 8, cout, dt)
    cb = torch.zeros(cout, device=dev)
    cases.append(lambda xc=xc, wf=wf, cb=cb, cd16=cd16, B=B, cout=cout:
                 ops.fused_conv2d(xc, wf, cb, cd16, B, 0, 8, 1, 0, None, 1.0, 8, cout, 3, 3, 1, 1, False, dt, True,
                                  drop_p=0.1, seed=seed))
    cases.append(lambda gyc=gyc, wb=wb, bt16=bt16, A=A, cin=cin:
                 ops.fused_conv2d(gyc, wb, None, bt16, A.contiguous(), 8, 9, cin * 9, -1, None, 1.0, 8, cin, 3, 3, 1, 1,
                                  True, dt, True))
# wgrad pair, dropout_dt
x, w, bias, d16, b, r = lin_case(4096, 320, 320)
gy = torch.randn(4096, 320, device=dev, dtype=dt)
T = torch.randn(4096, 16, device=dev)
dA, dB = torch.zeros(r, 320, device=dev), torch.zeros(320, r, device=dev)
cases.append(lambda: ops.wgrad_pair(x, T, dA, gy, T, dB, None, 1.0, r))
cases.append(lambda: ops.dropout_dt(gy, b, r, 1, 0.1, seed, r))
# one-launch optimizer step over an SD1.5-sized arena (192 sites, r = 4)
sites = nn.ModuleList([L.LoraInjectedLinear(1280, 1280, r=4, dropout_p=0.0) for _ in range(150)]).to(dev)
arena = LoraArena([(sites, 1e-4)])
arena.g.normal_()
cases.append(lambda: arena.step())
# SVD distillation (all kernels of lb_svd_truncated_batched)
from lora_b200.svd import svd_lowrank_ragged
Wb = [(torch.randn(1280, 1280, device=dev) * 0.05).half() for _ in range(8)]
Wt = [(w.float() + torch.randn(1280, 8, device=dev) @ torch.randn(8, 1280, device=dev) * 5e-4).half() for w in Wb]
cases.append(lambda: svd_lowrank_ragged(Wt, Wb, 8, clamp_quantile=0.99))
# step glue
from lora_b200.host.ddpm import DDPMNoiser
from lora_b200.step_ops import fused_masked_mse, step_prologue
noiser = DDPMNoiser(device=dev)
lat, eps = torch.randn(1, 4, 64, 64, device=dev), torch.randn(1, 4, 64, 64, device=dev)
tt = torch.tensor([500], device=dev)
pred = torch.randn(1, 4, 64, 64, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
cases.append(lambda: step_prologue(lat, eps, tt, noiser, dt))
cases.append(lambda: fused_masked_mse(pred, eps))

for f in cases:          # warm-up (allocates the split-K workspace, instantiates kernels)
    f(); f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for f in cases:
    f()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", len(cases), "cases")
