#!/bin/bash
TAG=${1:-r03c}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_paired.py tests/test_gpu_headline.py -x -q -m gpu -k "paired or propagation" > $OUT/${TAG}_pytest.log 2>&1
timeout 200 python tools/stab_breakdown.py 16 > $OUT/${TAG}_stab.log 2>&1
timeout 200 python tools/qr_profile_paired.py 16 > $OUT/${TAG}_qr_profile_paired.log 2>&1
tail -3 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_stab.log;  head -20 $OUT/${TAG}_qr_profile_paired.log
