#!/bin/bash
# Multi-GPU visit: data-parallel bench at N GPUs (torchrun, NCCL) + DP parity test.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_dp.txt 2>&1
echo "=== bench N=$N" > gpurun_out/dp.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps ${DP_STEPS:-20} --warmup 3 > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
tail -5 gpurun_out/bench_dp$N.err >> gpurun_out/dp.log; cat gpurun_out/bench_dp$N.json >> gpurun_out/dp.log
if [ -n "$DP_FULL" ]; then
echo "=== reference arm under torchrun" >> gpurun_out/dp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 >> gpurun_out/dp.log 2>&1
echo "=== dp parity test" >> gpurun_out/dp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  scripts/dp_parity.py >> gpurun_out/dp.log 2>&1
fi
tail -20 gpurun_out/dp.log
