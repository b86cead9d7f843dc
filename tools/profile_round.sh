#!/bin/bash
# ncu evidence for one round (run under gpurun):  bash tools/profile_round.sh <tag>
# 1. launch list of a short bench run (shares of the sweep), 2. --set full captures of the top kernels (one launch group each).
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
# launch list of one steady-state sweep (init ~5.8 k launches + one warm-up sweep ~12 k are skipped; graph kernel nodes count as launches)
ncu --metrics gpu__time_duration.sum --clock-control none -s 18000 -c 12500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --two-chains 0 --bfield-series 0 --extra-configs 0 > $out/${tag}_ncu_bench.log 2>&1
for spec in "lu:lu_block:5:1:1" "panel:qr_panel:3:3" "larfb:larfb_kernel:3:3" "larfbc:larfb_cluster:3:3" "zgemm:zgemm_kernel:1:2" "wrap:apply_chain:0:6:14" "trsm:trsm_kernel:4:1" "qrcp:qrcp_kernel:99:1"; do
  IFS=: read name regex which cnt skip <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$regex -s ${skip:-0} -c $cnt -f -o $out/${tag}_$name \
      python tools/prof_target.py 16 $which > $out/${tag}_ncu_$name.log 2>&1
done
ls -la $out | grep ${tag}_
