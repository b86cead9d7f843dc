#!/bin/bash
# One GPU validation pass (run through gpurun): local-update kernel timing/profile, the GPU test-suite, a short bench.
# usage: tools/gpu_round.sh <tag> [pytest-args...]
TAG=${1:-r02}; shift
OUT=gpurun_out
mkdir -p $OUT
{
  echo "== lu_profile block kernel"; timeout 300 python tools/lu_profile.py 16
  echo "== lu_profile site kernel"; DQMC_LU_KERNEL=site timeout 300 python tools/lu_profile.py 16
} > $OUT/${TAG}_lu_profile.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q -s --durations=15 -p no:cacheprovider --timeout=400 --timeout-method=thread "$@" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
timeout 1200 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
tail -4 $OUT/${TAG}_lu_profile.log
tail -6 $OUT/${TAG}_pytest.log
