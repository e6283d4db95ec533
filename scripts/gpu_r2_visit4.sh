#!/bin/bash
# Round-2 visit 4 (1 GPU): deferred/batched wgrad, optimizer shadow phase fix, rank chunks, conv roofline;
# whole suite, C2 / C3 bench lines (with the reference-on-this-GPU block for C2), SVD launch list,
# ncu --set full of every kernel.
mkdir -p gpurun_out
L=gpurun_out/v4.log
: > $L
for f in tests/test_modules_gpu.py tests/test_ti_gpu.py tests/test_grouping_gpu.py tests/test_dropout_gpu.py tests/test_conv_gpu.py; do
  echo "=== $f" >> $L
  timeout 900 python -m pytest $f -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -12 >> $L
done
echo "=== pytest -m gpu (all)" >> $L
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -8 >> $L
echo "=== bench native C2 (full line)" >> $L
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/v4_bench.json 2>> $L
cat gpurun_out/v4_bench.json >> $L
echo "=== bench native C2, immediate wgrad" >> $L
LB_NO_DEFER=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v4_bench_nodefer.json 2>> $L
cat gpurun_out/v4_bench_nodefer.json >> $L
echo "=== bench extended (C3 shape, 1 GPU)" >> $L
timeout 900 python bench.py --extended --rank 8 --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v4_bench_ext.json 2>> $L
cat gpurun_out/v4_bench_ext.json >> $L
echo "=== bench PTI shape (C4: 768px rank 16, 1 GPU)" >> $L
timeout 900 python bench.py --res 768 --rank 16 --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v4_bench_pti.json 2>> $L
cat gpurun_out/v4_bench_pti.json >> $L
echo "=== svd launch list" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mul_right|mul_left|probes|jacobi32|tall_transform|factors_kernel|quantile' --csv \
   --log-file gpurun_out/v4_svd_launches.csv python scripts/bench_svd.py >> $L 2>&1
echo "=== ncu --set full, every kernel" >> $L
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f \
   -o gpurun_out/r2_kernels python scripts/prof_kernels_full.py >> $L 2>&1
ls -la gpurun_out/r2_kernels.ncu-rep >> $L 2>&1
tail -100 $L | cut -c1-1500
