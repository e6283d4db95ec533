#!/bin/bash
# First GPU visit of round 2: (1) the round-1 suite incl. the PTI-variant tests written after the
# round-1 budget ran out, (2) the experimental cluster split-K kernel in its OWN process (a trap
# there must not poison the rest), (3) its timings if it passed, (4) the bench line.
mkdir -p gpurun_out
L=gpurun_out/round2_first.log
echo "=== pytest -m gpu" > $L
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -25 >> $L
echo "=== experimental split-K parity" >> $L
LB_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_zz_splitk_experimental_gpu.py -x -q --timeout 120 -p no:cacheprovider 2>&1 | tail -25 >> $L
if tail -3 $L | grep -q " passed" && ! tail -3 $L | grep -q "failed"; then
  echo "=== split-K timings (warm)" >> $L
  timeout 600 python scripts/prof_splitk.py 2>&1 | tail -60 >> $L
fi
echo "=== bench" >> $L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_round2_first.json 2>> $L
cat gpurun_out/bench_round2_first.json >> $L
tail -40 $L
