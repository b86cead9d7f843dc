#!/bin/bash
TAG=${1:-r03c}
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_paired.py tests/test_gpu_headline.py tests/test_gpu_parity.py -x -q -m gpu -k "paired or propagation or inv_sum or calculate_greens or full_size" -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
timeout 200 python tools/stab_breakdown.py 16 > $OUT/${TAG}_stab.log 2>&1
tail -3 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_stab.log
