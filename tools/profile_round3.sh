#!/bin/bash
# ncu evidence for round 3 (run under gpurun, one part per call):  bash tools/profile_round3.sh <tag> list|full
# list: launch list of the timed sweep of `bench.py --steps 1 --warmup 1` (cudaProfilerStart/Stop around the timed region).
# full: --set full captures of the kernels of the half-matrix stabilization path, the local updates and the wrap; the reports stay
#       on the box (/tmp), only the one-line-per-launch table comes back (gpurun_out is limited to 64 MiB).
tag=${1:-r03}; part=${2:-list}
out=gpurun_out
mkdir -p $out
if [ "$part" = list ]; then
  DQMC_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --two-chains 0 --bfield-series 0 --extra-configs 0 \
      > $out/${tag}_ncu_bench.log 2>&1
  python tools/launch_summary.py $out/${tag}_launches.csv > $out/${tag}_launch_summary.md 2>&1
  head -40 $out/${tag}_launch_summary.md
else
  reps=""
  for spec in "lu:lu_block:5:1:1" "ppanel:qr_panel_paired:15:2:2" "narrow:larfb_narrow:15:2:2" "larfb:larfb_kernel:15:2:2" "wrap:apply_chain:0:4:14" "trsm:trsm_kernel:4:1"; do
    IFS=: read name regex which cnt skip <<< "$spec"
    timeout 150 ncu --set full --clock-control none -k regex:$regex -s ${skip:-0} -c $cnt -f -o /tmp/${tag}_$name \
        python tools/prof_target.py 16 $which > $out/${tag}_ncu_$name.log 2>&1
    [ -f /tmp/${tag}_$name.ncu-rep ] && reps="$reps /tmp/${tag}_$name.ncu-rep"
  done
  python tools/ncu_summary.py $reps > $out/${tag}_ncu_table.md 2>&1
  cat $out/${tag}_ncu_table.md
fi
