"""`ncu -i report.ncu-rep --page raw --csv` -> one markdown row per kernel with the metrics the
roofline discussion uses (duration, DRAM bytes and throughput, tensor-pipe activity, warps active,
registers, shared memory). Usage: ncu -i X.ncu-rep --page raw --csv | python scripts/summarize_ncu_full.py"""
import csv
import re
import sys

rd = csv.reader(sys.stdin)
rows = [r for r in rd if r]
hdr = rows[0]
units = rows[1]
data = rows[2:]
col = {name: i for i, name in enumerate(hdr)}


def get(r, name, default=""):
    i = col.get(name)
    return r[i] if i is not None and i < len(r) else default


want = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
print("| # | kernel | " + " | ".join(lbl for _, lbl in want) + " |")
print("|---|---|" + "---|" * len(want))
for n, r in enumerate(data):
    name = re.sub(r"\(.*", "", get(r, "Kernel Name"))[-70:]
    cells = []
    for key, _ in want:
        v = get(r, key)
        u = units[col[key]] if key in col else ""
        cells.append(f"{v} {u}".strip())
    print(f"| {n} | {name} | " + " | ".join(cells) + " |")
