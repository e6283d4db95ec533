"""Per-site device times of the fused LoRA-linear kernel against cuBLAS's base GEMM alone, from
ncu's gpu__time_duration (NOT Python event pairs: these kernels are a few microseconds long).

Run (one GPU, under gpurun):
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
        --csv --log-file gpurun_out/sites_ncu.csv python scripts/prof_sites_ncu.py
    python scripts/prof_sites_ncu.py --summarize gpurun_out/sites_ncu.csv gpurun_out/sites_plan.json \
        > profiles/r2_site_table.md

For each row of SURVEY.md Appendix A (M, K, N) x {forward, dX = the same kernel on (M, N, K)} the
profiled section launches, REPS times each:  marker | ours (X.W^T + bias + LoRA, T side output) |
marker | torch.matmul(X, W^T) (base GEMM only, what cuBLAS picks) ; a marker is a 1-element int64
add -- the only elementwise kernel in the section -- so the summariser can split the launch list
without knowing kernel names. ncu flushes caches before every launch (cold L2, like a site inside
a step whose working set exceeds L2); the reported figure is the MEDIAN over REPS.
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (label, M, K, N, r)  -- SURVEY.md Appendix A (512x512, r = 4) + the CLIP row
SITES = [
    ("L0 attn q/k/v/o", 4096, 320, 320, 4), ("L0 attn2 k/v", 77, 768, 320, 4), ("L0 GEGLU", 4096, 320, 2560, 4),
    ("L1 attn q/k/v/o", 1024, 640, 640, 4), ("L1 attn2 k/v", 77, 768, 640, 4), ("L1 GEGLU", 1024, 640, 5120, 4),
    ("L2 attn q/k/v/o", 256, 1280, 1280, 4), ("L2 attn2 k/v", 77, 768, 1280, 4), ("L2 GEGLU", 256, 1280, 10240, 4),
    ("mid attn q/k/v/o", 64, 1280, 1280, 4), ("mid GEGLU", 64, 1280, 10240, 4), ("CLIP k/v/q/out", 77, 768, 768, 4),
]
REPS = int(os.environ.get("REPS", 5))
if os.environ.get("SITES") == "small":          # the <= 148-tile sites (split-K planner experiments)
    SITES = [s for s in SITES if s[1] <= 256]
if os.environ.get("SITES") == "geglu":          # the wide projections + the mid block (schedule experiments)
    SITES = [s for s in SITES if "GEGLU" in s[0] or s[0].startswith("mid")]
TAG = os.environ.get("TAG", "")
TILED = os.environ.get("TILED") == "1"          # frozen weight in the 64x64-block layout (LB_W_TILED)
MODE = int(os.environ.get("MODE", 0))           # lb_debug_set_linear_mode value (0 = the planner's choice)


def run():
    import torch
    from lora_b200 import ops
    dev, dt = "cuda", torch.bfloat16
    if MODE:
        from lora_b200 import _C
        assert _C.lib.lb_debug_set_linear_mode(MODE) == 0
    marker = torch.zeros(1, device=dev, dtype=torch.int64)
    plan = []
    cases = []
    for (label, M, K, N, r) in SITES:
        for direction, (m, k, n) in (("fwd", (M, K, N)), ("dX", (M, N, K))):
            x = torch.randn(m, k, device=dev, dtype=dt)
            w = torch.randn(n, k, device=dev, dtype=dt) * 0.02
            a = torch.randn(r, k, device=dev)
            b = torch.randn(n, r, device=dev) * 0.01
            d16 = ops.cast_rows_pad16(a, k, 1, r, k, dt)
            bias = torch.zeros(n, device=dev) if direction == "fwd" else None
            wt = w.t()
            wo = ops.tile_weight(w, dt) if TILED else w
            ours = lambda x=x, w=wo, bias=bias, d16=d16, b=b, r=r: ops.fused_linear(x, w, bias, d16, b, r, 1, None, 1.0, r, dt, True)
            lib = lambda x=x, wt=wt: torch.matmul(x, wt)
            for _ in range(3):
                ours(); lib()
            cases.append((label, direction, m, k, n, r, ours, lib))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for (label, direction, m, k, n, r, ours, lib) in cases:
        for rep in range(REPS):
            marker.add_(1); ours()
            marker.add_(1); lib()
        plan.append({"label": label, "dir": direction, "M": m, "K": k, "N": n, "r": r, "reps": REPS})
    marker.add_(1)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(plan, open(os.path.join(ROOT, "gpurun_out", f"sites_plan{TAG}.json"), "w"), indent=1)


def summarize(csv_path, plan_path):
    import csv
    with open(csv_path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    segs, cur = [], None
    for row in rd:
        name, ns = row[ik], float(row[iv].replace(",", ""))
        if "elementwise" in name:
            if cur is not None:
                segs.append(cur)
            cur = []
        elif cur is not None:
            cur.append((name, ns))
    plan = json.load(open(plan_path))
    print("| site | dir | M | K | N | fused LoRA kernel us | cuBLAS base GEMM us | ratio ours/cuBLAS | kernel(s) |")
    print("|---|---|---|---|---|---|---|---|---|")
    i, wins, total = 0, 0, 0
    for p in plan:
        ours, lib, names = [], [], set()
        for _ in range(p["reps"]):
            ours.append(sum(ns for _, ns in segs[i]) / 1e3)
            names.update(n.split("(")[0].split("<")[0][-40:] for n, _ in segs[i])
            lib.append(sum(ns for _, ns in segs[i + 1]) / 1e3)
            libn = len(segs[i + 1])
            i += 2
        o, l = statistics.median(ours), statistics.median(lib)
        total += 1
        wins += o <= l
        print(f"| {p['label']} | {p['dir']} | {p['M']} | {p['K']} | {p['N']} | {o:.2f} | {l:.2f} ({libn} launch) | {o / l:.2f} | {', '.join(sorted(names))} |")
    print(f"\nfused <= cuBLAS base GEMM on {wins} of {total} (site, direction) rows. ncu gpu__time_duration, cold L2 "
          f"(ncu flushes caches per launch), median of {plan[0]['reps']}; the fused kernel ALSO computes the rank-r "
          "branch, adds the bias and writes T [M,16] fp32 -- the cuBLAS column is the frozen GEMM alone.")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--summarize":
        summarize(sys.argv[2], sys.argv[3])
    else:
        run()
