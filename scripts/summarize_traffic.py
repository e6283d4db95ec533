"""DRAM traffic of the fused-kernel sweep from an ncu CSV
(ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:fused_lora --csv
 python bench.py --roofline-only): sums the LAST `launches` fused launches -> profiles/fused_linear_dram_traffic.json"""
import csv, json, sys
path, n = sys.argv[1], int(sys.argv[2])
lines = [l for l in open(path) if l.startswith('"')]
rd = csv.reader(lines); hdr = next(rd)
iid, im, iv, iu = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = {}
for r in rd:
    v = float(r[iv].replace(",", ""))
    u = r[iu].lower()
    if "byte" in u:
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    per.setdefault(int(r[iid]), {})[r[im]] = v
ids = sorted(per)[-n:]
rd_b = sum(per[i].get("dram__bytes_read.sum", 0) for i in ids)
wr_b = sum(per[i].get("dram__bytes_write.sum", 0) for i in ids)
t_ns = sum(per[i].get("gpu__time_duration.sum", 0) for i in ids)
out = {"launches": len(ids), "dram_bytes_read": rd_b, "dram_bytes_write": wr_b,
       "dram_bytes_per_sweep": rd_b + wr_b, "sum_kernel_time_ns_under_ncu": t_ns,
       "note": "one pass per launch under ncu: L2 is NOT flushed between launches of the sweep, outputs that "
               "fit the 126 MB L2 are written back later and show up as 0 write bytes on their own launch"}
print(json.dumps(out))
json.dump(out, open("profiles/fused_linear_dram_traffic.json", "w"), indent=1)
