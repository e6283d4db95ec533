#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest subset" > gpurun_out/quick.log
timeout 900 python -m pytest tests/test_grouping_gpu.py tests/test_ti_gpu.py tests/test_svd_gpu.py tests/test_modules_gpu.py tests/test_kernels_gpu.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -60 >> gpurun_out/quick.log
echo "=== bench (grouped)" >> gpurun_out/quick.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_grouped.json 2>> gpurun_out/quick.log
cat gpurun_out/bench_grouped.json >> gpurun_out/quick.log
echo "=== bench (ungrouped)" >> gpurun_out/quick.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-group > gpurun_out/bench_ungrouped.json 2>> gpurun_out/quick.log
cat gpurun_out/bench_ungrouped.json >> gpurun_out/quick.log
tail -30 gpurun_out/quick.log
