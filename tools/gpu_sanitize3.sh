#!/bin/bash
# compute-sanitizer on the round-3 kernels (L=4: n=64, two paired panels per factorization; a few updates through propagate)
OUT=gpurun_out; mkdir -p $OUT
F=$OUT/r03_sanitizer.txt
echo "# compute-sanitizer on tools/sanitize_small.py (L=4, M=20, SAN_UPDATES=22: init + two stabilizations through the half-matrix path), round 3" > $F
echo "## memcheck (whole run, no TDGF)" >> $F
SAN_TDGF=0 SAN_UPDATES=22 timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -4 >> $F
for k in qr_panel_paired larfb_narrow apply_chain; do
  echo "## racecheck --kernel-regex kns=$k" >> $F
  SAN_TDGF=0 SAN_UPDATES=12 timeout 150 compute-sanitizer --tool racecheck --kernel-regex kns=$k python tools/sanitize_small.py 2>&1 | tail -3 >> $F
done
cat $F
