#!/bin/bash
# quick validation: headline parity + paired tests + sweep parity at small sizes, then the short bench
TAG=${1:-r03q}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_paired.py tests/test_gpu_headline.py tests/test_gpu_parity.py -x -q -m gpu -k "paired or propagation or local_updates or shared_stream or wrap or symmetry or flavour" -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
tail -4 $OUT/${TAG}_pytest.log
bash tools/gpu_bench.sh $TAG
