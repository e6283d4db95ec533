"""Launch the fused LoRA-linear kernel on representative SD1.5 site shapes (for ncu --set full)
and print CUDA-event timings per shape (algorithmic GB/s, TFLOP/s)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lora_b200 import ops

SHAPES = [(4096, 320, 2560, 4), (4096, 320, 320, 4), (1024, 640, 640, 4), (1024, 640, 5120, 4),
          (256, 1280, 1280, 4), (256, 1280, 10240, 4), (64, 1280, 10240, 4), (77, 768, 768, 4),
          (77, 768, 320, 4), (9216, 320, 2560, 16)]
dev = "cuda"
dt = torch.bfloat16
rows = []
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for (M, K, N, r) in SHAPES:
    x = torch.randn(M, K, device=dev, dtype=dt)
    w = torch.randn(N, K, device=dev, dtype=dt) * 0.02
    a = torch.randn(r, K, device=dev); b = torch.randn(N, r, device=dev) * 0.01
    d16 = ops.cast_rows_pad16(a, K, 1, r, K, dt)
    bias = torch.zeros(N, device=dev)
    for _ in range(3):
        ops.fused_linear(x, w, bias, d16, b, r, 1, None, 1.0, r, dt, True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()                       # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.fused_linear(x, w, bias, d16, b, r, 1, None, 1.0, r, dt, True); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    byts = 2 * (M * K + N * K + r * K + M * N) + 4 * N * r + 4 * N + 4 * M * 16
    fl = 2 * M * K * N + 2 * M * r * (K + N)
    rows.append({"M": M, "K": K, "N": N, "r": r, "us": ms * 1e3, "GBps": byts / ms / 1e6, "TFLOPs": fl / ms / 1e9})
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/site_times.json", "w"), indent=1)
