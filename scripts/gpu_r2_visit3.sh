#!/bin/bash
# Round-2 visit 3: split-K v2 (L2 vector reductions), one-launch optimizer step, step prologue /
# loss epilogue kernels, TI loop, dropout_dt rewrite. Focused test files first, then the suite,
# site table, C2 (split on/off) and C3 bench lines, launch lists of the SVD bench and one C3 step.
mkdir -p gpurun_out
L=gpurun_out/v3.log
: > $L
for f in tests/test_splitk_gpu.py tests/test_conv_gpu.py tests/test_dropout_gpu.py tests/test_ti_gpu.py tests/test_modules_gpu.py tests/test_zz_pti_variants_gpu.py; do
  echo "=== $f" >> $L
  timeout 900 python -m pytest $f -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -25 >> $L
done
echo "=== pytest -m gpu (all)" >> $L
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -12 >> $L
echo "=== site table (ncu)" >> $L
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/sites_ncu3.csv python scripts/prof_sites_ncu.py >> $L 2>&1
python scripts/prof_sites_ncu.py --summarize gpurun_out/sites_ncu3.csv gpurun_out/sites_plan.json > gpurun_out/site_table3.md 2>> $L
cat gpurun_out/site_table3.md >> $L
echo "=== bench native C2" >> $L
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v3_bench.json 2>> $L
cat gpurun_out/v3_bench.json >> $L
echo "=== bench native C2, no split-K" >> $L
LB_NO_SPLITK=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v3_bench_nosplit.json 2>> $L
cat gpurun_out/v3_bench_nosplit.json >> $L
echo "=== bench extended (C3 shape, 1 GPU)" >> $L
timeout 900 python bench.py --extended --rank 8 --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-baseline > gpurun_out/v3_bench_ext.json 2>> $L
cat gpurun_out/v3_bench_ext.json >> $L
echo "=== bench svd + launch list" >> $L
timeout 600 python scripts/bench_svd.py >> $L 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lbsvd2 --csv \
   --log-file gpurun_out/v3_svd_launches.csv python scripts/bench_svd.py >> /dev/null 2>&1
echo "=== launch list: one extended step" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/v3_ext_launches.csv python bench.py --extended --rank 8 --profile-steps 1 >> $L 2>&1
python scripts/summarize_launches.py gpurun_out/v3_ext_launches.csv >> $L 2>&1
tail -120 $L
