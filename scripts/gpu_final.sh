#!/bin/bash
# Evidence visit: full GPU test-suite, smoke, bench (+reference arm), other configs, ncu evidence.
mkdir -p gpurun_out
L=gpurun_out/final.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest gpu" > $L
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -40 >> $L
echo "=== smoke" >> $L
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3 >> $L
echo "=== bench" >> $L
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err >> $L; cat gpurun_out/bench.json >> $L
echo "=== bench reference arm" >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> $L
cat gpurun_out/bench_ref.json >> $L
if [ -n "$DO_CONFIGS" ]; then
echo "=== bench extended (configs[2] shape, 1 GPU)" >> $L
timeout 900 python bench.py --extended --rank 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_extended.json 2>> $L
cat gpurun_out/bench_extended.json >> $L
echo "=== bench PTI shape (configs[3]: r16 768px, 1 GPU)" >> $L
timeout 900 python bench.py --res 768 --rank 16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pti.json 2>> $L
cat gpurun_out/bench_pti.json >> $L
echo "=== bench svd (configs[4])" >> $L
timeout 600 python scripts/bench_svd.py >> $L 2>&1
fi
echo "=== ncu launch list" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-steps 2 --no-graph --no-cpu-baseline >> $L 2>&1
echo "=== ncu dram traffic of the fused sweep" >> $L
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fused_lora --csv --log-file gpurun_out/traffic.csv python bench.py --roofline-only --no-graph --no-cpu-baseline >> $L 2>&1
tail -30 $L
