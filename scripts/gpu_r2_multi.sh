#!/bin/bash
# Multi-GPU visit (N = $1 GPUs of one box; configs in $2 = space-separated list of c2 | c3 | c4):
# data-parallel parity with the gradient exchange fused into the optimizer launch (NVLink peer reads,
# lb_optim_step_dp) and with NCCL, then bench lines per BASELINE config: the fused exchange (one
# graph per step) always, the NCCL exchange (two graphs) when $3 = nccl.
N=${1:-2}
CFGS=${2:-c2}
WITH_NCCL=${3:-}
mkdir -p gpurun_out
L=gpurun_out/multi_n$N.log
nvidia-smi --query-gpu=index,name --format=csv > $L 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== dp parity (peer exchange)" >> $L
timeout 300 $TR --master-port 29513 scripts/dp_parity.py 2>&1 | grep -E "dp_parity|Error|error" >> $L
if [ -n "$WITH_NCCL" ]; then
  echo "=== dp parity (NCCL exchange)" >> $L
  LB_DP_NCCL=1 timeout 300 $TR --master-port 29514 scripts/dp_parity.py 2>&1 | grep -E "dp_parity|Error|error" >> $L
fi
PORT=29520
for CFG in $CFGS; do
  case $CFG in
    c2) ARGS="" ;;
    c3) ARGS="--extended --rank 8" ;;
    c4) ARGS="--res 768 --rank 16" ;;
  esac
  echo "=== bench $CFG N=$N, exchange fused into the optimizer launch (one graph per step)" >> $L
  PORT=$((PORT+1))
  timeout 600 $TR --master-port $PORT bench.py --gpus $N --steps 20 --warmup 3 $ARGS > gpurun_out/bench_${CFG}_dp${N}_peer.json 2>> $L
  cat gpurun_out/bench_${CFG}_dp${N}_peer.json >> $L
  if [ -n "$WITH_NCCL" ]; then
    echo "=== bench $CFG N=$N, NCCL all-reduce between two graphs" >> $L
    PORT=$((PORT+1))
    timeout 600 $TR --master-port $PORT bench.py --gpus $N --steps 20 --warmup 3 --nccl-allreduce $ARGS > gpurun_out/bench_${CFG}_dp${N}_nccl.json 2>> $L
    cat gpurun_out/bench_${CFG}_dp${N}_nccl.json >> $L
  fi
done
if [ "$N" = "2" ]; then
  echo "=== pytest multi-GPU tests" >> $L
  timeout 600 python -m pytest tests/test_dp_gpu.py -q --timeout 600 -p no:cacheprovider 2>&1 | tail -5 >> $L
fi
grep -v "Warning\|Consider\|return float\|OMP_NUM\|\*\*\*\*" $L | tail -40 | cut -c1-1200
