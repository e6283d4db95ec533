#!/bin/bash
# One GPU-box visit: tests, smoke, bench, per-site timings, ncu launch list (+ optional full capture).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest gpu" > gpurun_out/round.log
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -80 >> gpurun_out/round.log
echo "=== smoke" >> gpurun_out/round.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -5 >> gpurun_out/round.log
echo "=== bench" >> gpurun_out/round.log
timeout 1200 python bench.py --steps 20 --warmup 3 ${BENCH_FLAGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err >> gpurun_out/round.log; cat gpurun_out/bench.json >> gpurun_out/round.log
echo "=== site times" >> gpurun_out/round.log
timeout 300 python scripts/prof_site.py >> gpurun_out/round.log 2>&1
if [ -n "$DO_LAUNCHES" ]; then
echo "=== ncu launch list" >> gpurun_out/round.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-steps 1 --no-graph --no-cpu-baseline >> gpurun_out/round.log 2>&1
fi
if [ -n "$DO_NCU_FULL" ]; then
echo "=== ncu full" >> gpurun_out/round.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_lora -c 12 -o gpurun_out/prof_fused -f python scripts/prof_site.py >> gpurun_out/round.log 2>&1
fi
tail -30 gpurun_out/round.log
