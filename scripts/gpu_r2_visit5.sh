#!/bin/bash
# Round-2 visit 5 (1 GPU): SVD passes with the cp.async ring + adaptive Jacobi, grouping stale check,
# new step-op tests; suite; SVD bench + launch list; smoke().
mkdir -p gpurun_out
L=gpurun_out/v5.log
: > $L
for f in tests/test_svd_gpu.py tests/test_grouping_gpu.py tests/test_step_ops_gpu.py; do
  echo "=== $f" >> $L
  timeout 900 python -m pytest $f -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -12 >> $L
done
echo "=== pytest -m gpu (all)" >> $L
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -8 >> $L
echo "=== bench svd" >> $L
timeout 600 python scripts/bench_svd.py >> $L 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mul_right|mul_left|probes|jacobi32|tall_transform|factors_kernel|quantile' --csv \
   --log-file gpurun_out/v5_svd_launches.csv python scripts/bench_svd.py > /dev/null 2>&1
echo "=== smoke()" >> $L
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "=== bench --impl reference-cuda (own line)" >> $L
timeout 900 python bench.py --impl reference-cuda --steps 20 --warmup 3 > gpurun_out/v5_bench_refcuda.json 2>> $L
cut -c1-600 gpurun_out/v5_bench_refcuda.json >> $L
grep -v "Warning\|Consider\|^$\|importlib\|swigvar\|-- Docs" $L | tail -60 | cut -c1-1000
