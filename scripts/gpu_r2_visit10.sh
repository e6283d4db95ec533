#!/bin/bash
# Round-2 visit 10 (1 GPU): dA/dB batches overlapped with backward on a side stream (LB_WGRAD_OVERLAP=1),
# same-box A/B on C2 and C3; the new test.
mkdir -p gpurun_out
L=gpurun_out/v10.log
echo "build $(cut -c1-12 lora_b200/.liblora_b200.stamp)" > $L
echo "=== tests" >> $L
timeout 900 python -m pytest tests/test_modules_gpu.py -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -5 >> $L
for ov in 0 1 0 1; do
  echo "=== bench C2 LB_WGRAD_OVERLAP=$ov" >> $L
  LB_WGRAD_OVERLAP=$ov timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v10_bench_ov$ov.json 2>> $L
  cut -c1-290 gpurun_out/v10_bench_ov$ov.json >> $L
done
for ov in 0 1; do
  echo "=== bench C3 LB_WGRAD_OVERLAP=$ov" >> $L
  LB_WGRAD_OVERLAP=$ov timeout 900 python bench.py --extended --rank 8 --steps 40 --warmup 5 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v10_bench_ext_ov$ov.json 2>> $L
  cut -c1-290 gpurun_out/v10_bench_ext_ov$ov.json >> $L
done
grep -v "Warning\|Consider\|^$\|importlib\|swigvar\|-- Docs" $L | tail -40 | cut -c1-300
