#!/bin/bash
# ncu --set full capture (with source) of one launch of the local-update kernel at the headline size; run through gpurun
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lu_block -s 1 -c 1 -f -o $OUT/${TAG}_lu_block \
    python tools/prof_target.py 16 5 > $OUT/${TAG}_ncu_lu.log 2>&1
ls -la $OUT | grep ${TAG}_
