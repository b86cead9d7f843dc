#!/bin/bash
# ncu evidence for round 3 (run under gpurun):  bash tools/profile_round3.sh <tag>
# 1. launch list of one steady-state sweep, 2. --set full captures of the kernels of the half-matrix stabilization path and the local updates.
tag=${1:-r03}
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -s 14000 -c 9000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --two-chains 0 --bfield-series 0 --extra-configs 0 > $out/${tag}_ncu_bench.log 2>&1
for spec in "lu:lu_block:5:1:1" "ppanel:qr_panel_paired:15:3:2" "narrow:larfb_narrow:15:3:2" "larfb:larfb_kernel:15:3:2" "zgemm:zgemm_kernel:4:2" "wrap:apply_chain:0:6:14" "trsm:trsm_kernel:4:1"; do
  IFS=: read name regex which cnt skip <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -s ${skip:-0} -c $cnt -f -o $out/${tag}_$name \
      python tools/prof_target.py 16 $which > $out/${tag}_ncu_$name.log 2>&1
done
ls -la $out | grep ${tag}_
