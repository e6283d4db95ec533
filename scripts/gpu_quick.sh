#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest subset" > gpurun_out/quick.log
timeout 900 python -m pytest tests/test_svd_gpu.py tests/test_dropout_gpu.py tests/test_modules_gpu.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -40 >> gpurun_out/quick.log
echo "=== site times" >> gpurun_out/quick.log
PROF_AUTO_ONLY=1 timeout 300 python scripts/prof_site.py >> gpurun_out/quick.log 2>&1
tail -40 gpurun_out/quick.log
