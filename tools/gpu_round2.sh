#!/bin/bash
# second-style GPU pass: ncu capture of the block kernel, its cycle profile, selected tests
TAG=${1:-r02}; shift
OUT=gpurun_out; mkdir -p $OUT
tools/gpu_ncu_lu.sh $TAG
{ echo "== lu_profile block kernel"; timeout 300 python tools/lu_profile.py 16; } > $OUT/${TAG}_lu_profile.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q -s --durations=10 -p no:cacheprovider --timeout=600 --timeout-method=thread "$@" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
cat $OUT/${TAG}_lu_profile.log; tail -12 $OUT/${TAG}_pytest.log
