#!/bin/bash
# Round-2 visit 9 (1 GPU): validation of the new planner (cluster split-K / 192-wide wave / narrow rule) as
# the driver would run things: whole GPU suite, smoke(), default bench (C2 + roofline + reference blocks),
# C3 / C4 lines, launch lists (sweep + one step), SVD launch list.
mkdir -p gpurun_out
L=gpurun_out/v9.log
echo "build $(cut -c1-12 lora_b200/.liblora_b200.stamp)" > $L
echo "=== pytest -m gpu (all)" >> $L
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -8 >> $L
echo "=== smoke()" >> $L
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> $L
echo "=== bench default (C2)" >> $L
timeout 1200 python bench.py > gpurun_out/v9_bench.json 2>> $L
cat gpurun_out/v9_bench.json >> $L
echo "=== bench C2 without the cluster split / any split (same box A/B)" >> $L
LB_NO_SPLITK=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v9_bench_nosplit.json 2>> $L
cut -c1-300 gpurun_out/v9_bench_nosplit.json >> $L
echo "=== bench extended (C3 shape, 1 GPU)" >> $L
timeout 900 python bench.py --extended --rank 8 --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v9_bench_ext.json 2>> $L
cat gpurun_out/v9_bench_ext.json >> $L
echo "=== bench PTI shape (C4: 768px rank 16, 1 GPU)" >> $L
timeout 900 python bench.py --res 768 --rank 16 --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v9_bench_pti.json 2>> $L
cat gpurun_out/v9_bench_pti.json >> $L
echo "=== sweep launch list (ncu)" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fused_lora' --csv \
   --log-file gpurun_out/v9_sweep_launches.csv python bench.py --roofline-only > /dev/null 2>> $L
echo "=== svd" >> $L
timeout 600 python scripts/bench_svd.py >> $L 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mul_right|mul_left|probes|jacobi32|tall_transform|factors_kernel|quantile' --csv \
   --log-file gpurun_out/v9_svd_launches.csv python scripts/bench_svd.py > /dev/null 2>&1
grep -v "Warning\|Consider\|^$\|importlib\|swigvar\|-- Docs" $L | tail -60 | cut -c1-1800
