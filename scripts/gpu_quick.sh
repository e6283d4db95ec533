#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest (all, -x)" > gpurun_out/quick.log
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -30 >> gpurun_out/quick.log
echo "=== bench (async wgrad)" >> gpurun_out/quick.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_async.json 2>> gpurun_out/quick.log
cat gpurun_out/bench_async.json >> gpurun_out/quick.log
tail -12 gpurun_out/quick.log
