#!/bin/bash
# ncu evidence for one round (run under gpurun):  bash tools/profile_round.sh <tag>
# 1. launch list of a short bench run (shares of the sweep), 2. --set full captures of the top kernels (one launch group each).
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -s 8000 -c 4000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 > $out/${tag}_ncu_bench.log 2>&1
for spec in "lu:local_updates:5:2" "panel:qr_panel:3:3" "larfb:larfb_kernel:3:3" "larfbc:larfb_cluster:3:3" "zgemm:zgemm_kernel:1:2" "wrap:apply_chain:0:6:14" "trsm:trsm_kernel:4:1"; do
  IFS=: read name regex which cnt skip <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$regex -s ${skip:-0} -c $cnt -f -o $out/${tag}_$name \
      python tools/prof_target.py 16 $which > $out/${tag}_ncu_$name.log 2>&1
done
ls -la $out | grep ${tag}_
