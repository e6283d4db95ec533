#!/bin/bash
# Multi-GPU visit (N = $1 GPUs of one box): data-parallel parity with the gradient exchange fused
# into the optimizer launch (NVLink peer reads, lb_optim_step_dp) and with NCCL; bench lines for the
# BASELINE config named in $2 (c2 | c3 | c4) with both exchanges. NCCL_DEBUG=INFO is left to the
# caller (bench.py honours an externally set value).
N=${1:-2}
CFG=${2:-c2}
mkdir -p gpurun_out
L=gpurun_out/multi_${CFG}_n$N.log
nvidia-smi --query-gpu=index,name --format=csv > $L 2>&1
nvidia-smi topo -m >> $L 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== dp parity (peer exchange)" >> $L
timeout 300 $TR --master-port 29513 scripts/dp_parity.py >> $L 2>&1
echo "=== dp parity (NCCL exchange)" >> $L
LB_DP_NCCL=1 timeout 300 $TR --master-port 29514 scripts/dp_parity.py >> $L 2>&1
case $CFG in
  c2) ARGS="" ;;
  c3) ARGS="--extended --rank 8" ;;
  c4) ARGS="--res 768 --rank 16" ;;
esac
echo "=== bench $CFG N=$N, exchange fused into the optimizer launch (one graph per step)" >> $L
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 $ARGS > gpurun_out/bench_${CFG}_dp${N}_peer.json 2>> $L
cat gpurun_out/bench_${CFG}_dp${N}_peer.json >> $L
echo "=== bench $CFG N=$N, NCCL all-reduce between two graphs" >> $L
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --nccl-allreduce $ARGS > gpurun_out/bench_${CFG}_dp${N}_nccl.json 2>> $L
cat gpurun_out/bench_${CFG}_dp${N}_nccl.json >> $L
if [ "$N" = "2" ]; then
  echo "=== pytest multi-GPU tests" >> $L
  timeout 600 python -m pytest tests/test_dp_gpu.py -q --timeout 600 -p no:cacheprovider 2>&1 | tail -5 >> $L
  echo "=== bench c2 N=1 on the same box" >> $L
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/bench_c2_dp1_samebox.json 2>> $L
  cat gpurun_out/bench_c2_dp1_samebox.json >> $L
fi
tail -60 $L | cut -c1-1200
