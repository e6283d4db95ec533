#!/bin/bash
# Round-2 visit 7 (1 GPU): same-box A/Bs. (a) dropout backward in-kernel mask vs two-pass on the extended
# step; (b) tight X box vs full box on the single-row-tile sites; (c) cluster split-K on the small sites;
# (d) tile schedules incl. BLOCK_N 192 on the wide projections; (e) SVD per-kernel list after the Jacobi
# and split-term changes.
mkdir -p gpurun_out
L=gpurun_out/v7.log
: > $L
for f in tests/test_kernels_gpu.py tests/test_svd_gpu.py; do
  echo "=== $f" >> $L
  timeout 900 python -m pytest $f -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -6 >> $L
done
echo "=== bench svd (default terms 2223, then 3333)" >> $L
timeout 600 python scripts/bench_svd.py >> $L 2>&1
LB_SVD_TERMS=3333 timeout 600 python scripts/bench_svd.py >> $L 2>&1
LB_SVD_TERMS=2123 timeout 600 python scripts/bench_svd.py >> $L 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mul_right|mul_left|probes|jacobi32|tall_transform|factors_kernel|quantile' --csv \
   --log-file gpurun_out/v7_svd_launches.csv python scripts/bench_svd.py > /dev/null 2>&1
prof() {  # tag, sites, extra env...
  local tag=$1 sites=$2; shift 2
  env "$@" SITES=$sites TAG=_$tag REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/sites_$tag.csv python scripts/prof_sites_ncu.py > /dev/null 2>&1
  python scripts/prof_sites_ncu.py --summarize gpurun_out/sites_$tag.csv gpurun_out/sites_plan_$tag.json > gpurun_out/site_table_$tag.md 2>> $L
  echo "=== sites $tag ($*)" >> $L
  cut -d'|' -f2-8 gpurun_out/site_table_$tag.md | head -26 >> $L
}
prof tight small LB_NO_OP=1
prof fullbox small LB_FULL_BOX=1
prof cl2 small MODE=$((3+4+16))
prof cl3 small MODE=$((3+4+32))
prof cl4 small MODE=$((3+4+48))
prof g_auto geglu LB_NO_OP=1
prof g_1t64 geglu MODE=5
prof g_1t128 geglu MODE=9
prof g_1t192 geglu MODE=13
prof g_p128 geglu MODE=10
echo "=== bench extended: in-kernel mask (default)" >> $L
timeout 900 python bench.py --extended --rank 8 --steps 20 --warmup 3 > gpurun_out/v7_bench_ext.json 2>> $L
cut -c1-330 gpurun_out/v7_bench_ext.json >> $L
echo "=== bench extended: two-pass dropout backward" >> $L
LB_DROPOUT_BWD=twopass timeout 900 python bench.py --extended --rank 8 --steps 20 --warmup 3 > gpurun_out/v7_bench_ext_twopass.json 2>> $L
cut -c1-330 gpurun_out/v7_bench_ext_twopass.json >> $L
echo "=== bench extended: in-kernel mask again" >> $L
timeout 900 python bench.py --extended --rank 8 --steps 20 --warmup 3 > gpurun_out/v7_bench_ext2.json 2>> $L
cut -c1-330 gpurun_out/v7_bench_ext2.json >> $L
grep -v "Warning\|Consider\|^$\|importlib\|swigvar\|-- Docs" $L | tail -200 | cut -c1-200
