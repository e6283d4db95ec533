#!/bin/bash
# Round-2 visit 6 (1 GPU): dropout backward with the in-kernel mask (BMASK), prologue fix, forced
# split-K sweep on the small sites (planner calibration), extended bench.
mkdir -p gpurun_out
L=gpurun_out/v6.log
: > $L
for f in tests/test_svd_gpu.py tests/test_dropout_gpu.py tests/test_step_ops_gpu.py tests/test_splitk_gpu.py; do
  echo "=== $f" >> $L
  timeout 900 python -m pytest $f -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -12 >> $L
done
echo "=== pytest -m gpu (all)" >> $L
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -8 >> $L
for F in 0 2 3 4 6; do
  echo "=== forced split $F (small sites)" >> $L
  SITES=small TAG=_f$F LB_SPLIT_FORCE=$F REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/sites_f$F.csv python scripts/prof_sites_ncu.py > /dev/null 2>&1
  python scripts/prof_sites_ncu.py --summarize gpurun_out/sites_f$F.csv gpurun_out/sites_plan_f$F.json > gpurun_out/site_table_f$F.md 2>> $L
  cut -d'|' -f2-9 gpurun_out/site_table_f$F.md | head -30 >> $L
done
echo "=== bench svd" >> $L
timeout 600 python scripts/bench_svd.py >> $L 2>&1
echo "=== bench extended" >> $L
timeout 900 python bench.py --extended --rank 8 --steps 20 --warmup 3 > gpurun_out/v6_bench_ext.json 2>> $L
cut -c1-400 gpurun_out/v6_bench_ext.json >> $L
echo "=== bench default" >> $L
timeout 900 python bench.py > gpurun_out/v6_bench.json 2>> $L
cut -c1-300 gpurun_out/v6_bench.json >> $L
grep -v "Warning\|Consider\|^$\|importlib\|swigvar\|-- Docs" $L | tail -150 | cut -c1-400
