#!/bin/bash
# Round-2 visit 1: whole GPU suite (incl. the new full-size step parity tests), the experimental
# cluster split-K kernel in its own process, per-site ncu table (baseline), bench (native with the
# reference-on-this-GPU block; PDL on/off), launch lists (fused sweep; one extended step).
mkdir -p gpurun_out
L=gpurun_out/v1.log
echo "=== pytest -m gpu" > $L
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -40 >> $L
echo "=== experimental split-K parity" >> $L
LB_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_splitk_experimental_gpu.py -x -q --timeout 120 -p no:cacheprovider 2>&1 | tail -15 >> $L
if tail -3 $L | grep -q " passed" && ! tail -3 $L | grep -q "failed"; then
  echo "=== split-K timings (warm)" >> $L
  REPS=100 timeout 300 python scripts/prof_splitk.py 2>&1 | tail -60 >> $L
fi
echo "=== site table (ncu)" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/sites_ncu.csv python scripts/prof_sites_ncu.py >> $L 2>&1
python scripts/prof_sites_ncu.py --summarize gpurun_out/sites_ncu.csv gpurun_out/sites_plan.json > gpurun_out/site_table.md 2>> $L
cat gpurun_out/site_table.md >> $L
echo "=== bench native" >> $L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/v1_bench.json 2>> $L
cat gpurun_out/v1_bench.json >> $L
echo "=== bench native, PDL on" >> $L
LB_PDL=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v1_bench_pdl.json 2>> $L
cat gpurun_out/v1_bench_pdl.json >> $L
echo "=== launch list: fused sweep" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fused_lora --csv \
   --log-file gpurun_out/v1_sweep_launches.csv python bench.py --roofline-only >> $L 2>&1
echo "=== launch list: one extended step" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/v1_ext_launches.csv python bench.py --extended --rank 8 --profile-steps 1 >> $L 2>&1
python scripts/summarize_launches.py gpurun_out/v1_ext_launches.csv >> $L 2>&1
tail -150 $L
