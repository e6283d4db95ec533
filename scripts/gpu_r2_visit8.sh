#!/bin/bash
# Round-2 visit 8 (1 GPU): frozen weights in the 64x64-block layout (LB_W_TILED) vs row-major, per site
# (ncu, cold L2); dropout backward A/B again with the one-stage-behind masked MMAs.
mkdir -p gpurun_out
L=gpurun_out/v8.log
echo "build $(cut -c1-12 lora_b200/.liblora_b200.stamp)" > $L
for f in tests/test_tiled_weight_gpu.py tests/test_dropout_gpu.py tests/test_kernels_gpu.py; do
  echo "=== $f" >> $L
  timeout 900 python -m pytest $f -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -6 >> $L
done
prof() {  # tag, sites, extra env...
  local tag=$1 sites=$2; shift 2
  env "$@" SITES=$sites TAG=_$tag REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/sites_$tag.csv python scripts/prof_sites_ncu.py > /dev/null 2>&1
  python scripts/prof_sites_ncu.py --summarize gpurun_out/sites_$tag.csv gpurun_out/sites_plan_$tag.json > gpurun_out/site_table_$tag.md 2>> $L
  echo "=== sites $tag ($*)" >> $L
  cut -d'|' -f2-8 gpurun_out/site_table_$tag.md | head -28 >> $L
}
prof rowmajor all TILED=0
prof tiled all TILED=1
prof tiled_1t192 geglu TILED=1 MODE=13
prof tiled_1t128 geglu TILED=1 MODE=9
echo "=== bench extended: two-pass dropout backward (default)" >> $L
timeout 900 python bench.py --extended --rank 8 --steps 20 --warmup 3 > gpurun_out/v8_bench_ext.json 2>> $L
cut -c1-330 gpurun_out/v8_bench_ext.json >> $L
echo "=== bench extended: mask inside the dX kernel, T MMAs one stage behind" >> $L
LB_DROPOUT_BWD=fused timeout 900 python bench.py --extended --rank 8 --steps 20 --warmup 3 > gpurun_out/v8_bench_ext_fused.json 2>> $L
cut -c1-330 gpurun_out/v8_bench_ext_fused.json >> $L
grep -v "Warning\|Consider\|^$\|importlib\|swigvar\|-- Docs" $L | tail -150 | cut -c1-200
