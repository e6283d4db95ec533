#!/bin/bash
# Round-2 visit 2: split-K plans, fused dropout drain, one-call SVD. Focused tests first (each file
# in its own process so that a trap in one kernel does not poison the rest), then the whole suite,
# the per-site ncu table, C2 / C3 bench lines and the SVD bench.
mkdir -p gpurun_out
L=gpurun_out/v2.log
: > $L
for f in tests/test_splitk_gpu.py tests/test_kernels_gpu.py tests/test_conv_gpu.py tests/test_dropout_gpu.py tests/test_svd_gpu.py; do
  echo "=== $f" >> $L
  timeout 600 python -m pytest $f -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -25 >> $L
done
echo "=== pytest -m gpu (all)" >> $L
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -15 >> $L
echo "=== experimental cluster split-K parity + timings" >> $L
LB_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_splitk_experimental_gpu.py -q --timeout 120 -p no:cacheprovider 2>&1 | tail -5 >> $L
REPS=100 timeout 300 python scripts/prof_splitk.py 2>&1 | tail -60 >> $L
echo "=== site table (ncu)" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/sites_ncu2.csv python scripts/prof_sites_ncu.py >> $L 2>&1
python scripts/prof_sites_ncu.py --summarize gpurun_out/sites_ncu2.csv gpurun_out/sites_plan.json > gpurun_out/site_table2.md 2>> $L
cat gpurun_out/site_table2.md >> $L
echo "=== bench native C2" >> $L
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v2_bench.json 2>> $L
cat gpurun_out/v2_bench.json >> $L
echo "=== bench native C2, no split-K" >> $L
LB_NO_SPLITK=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v2_bench_nosplit.json 2>> $L
cat gpurun_out/v2_bench_nosplit.json >> $L
echo "=== bench extended (C3 shape, 1 GPU)" >> $L
timeout 900 python bench.py --extended --rank 8 --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v2_bench_ext.json 2>> $L
cat gpurun_out/v2_bench_ext.json >> $L
echo "=== bench svd" >> $L
timeout 600 python scripts/bench_svd.py >> $L 2>&1
tail -230 $L
