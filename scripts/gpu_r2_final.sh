#!/bin/bash
# Round-2 final visit (1 GPU), as the driver runs things on a fresh box: whole GPU suite, smoke(), default
# bench and its reference arm; then the evidence passes for profiles/: DRAM traffic of the sweep, ncu
# --set full of every hand-written kernel.
mkdir -p gpurun_out
L=gpurun_out/final2.log
echo "build $(cut -c1-12 lora_b200/.liblora_b200.stamp)" > $L
echo "=== pytest -m gpu (all)" >> $L
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -6 >> $L
echo "=== smoke()" >> $L
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-200 >> $L
echo "=== bench default" >> $L
timeout 1200 python bench.py > gpurun_out/final2_bench.json 2>> $L
cut -c1-400 gpurun_out/final2_bench.json >> $L
echo "=== bench --impl reference (CPU port, bounded)" >> $L
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final2_bench_ref.json 2>> $L
cut -c1-400 gpurun_out/final2_bench_ref.json >> $L
for old in 1 0; do
  echo "=== bench extended, LB_CONV_NARROW_OLD=$old (conv planner A/B, same box)" >> $L
  LB_CONV_NARROW_OLD=$old timeout 900 python bench.py --extended --rank 8 --steps 30 --warmup 5 --no-cpu-baseline --no-cuda-baseline > gpurun_out/final2_bench_ext_old$old.json 2>> $L
  python - >> $L <<PY
import json
d = json.load(open("gpurun_out/final2_bench_ext_old$old.json"))
print("images/s %.2f  ms/step %.3f  conv sweep ms %.3f (frac %.4f)  linear sweep ms %.3f" % (d["value"], d["ms_per_step"], d["roofline_conv"]["ms_per_sweep"], d["roofline_conv"]["frac"], d["roofline"]["ms_per_sweep"]))
PY
done
echo "=== DRAM traffic of the sweep" >> $L
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'fused_lora' --csv \
   --log-file gpurun_out/final2_traffic.csv python bench.py --roofline-only > /dev/null 2>> $L
python scripts/summarize_traffic.py gpurun_out/final2_traffic.csv 240 >> $L 2>&1
cp profiles/fused_linear_dram_traffic.json gpurun_out/fused_linear_dram_traffic.json
echo "=== ncu --set full, every kernel" >> $L
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f \
   -o /tmp/r2_kernels_final python scripts/prof_kernels_full.py > /dev/null 2>&1
ls -la /tmp/r2_kernels_final.ncu-rep >> $L 2>&1
# the report itself is ~70 MB (gpurun_out/ is capped at 64 MiB): keep the summary and the raw CSV page
ncu -i /tmp/r2_kernels_final.ncu-rep --page raw --csv 2>/dev/null > /tmp/r2_raw.csv
python scripts/summarize_ncu_full.py < /tmp/r2_raw.csv > gpurun_out/r2_ncu_per_kernel_final.md 2>> $L
gzip -c /tmp/r2_raw.csv > gpurun_out/r2_ncu_raw_final.csv.gz
ls -la gpurun_out >> $L 2>&1
grep -v "Warning\|Consider\|^$\|importlib\|swigvar\|-- Docs" $L | tail -40 | cut -c1-600
