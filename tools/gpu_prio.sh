#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for p in 1 0; do
  DQMC_STREAM_PRIO=$p timeout 300 python bench.py --steps 4 --warmup 3 --two-chains 0 --bfield-series 0 --extra-configs 0 > $OUT/r03_prio$p.json 2> $OUT/r03_prio$p.err
  python - <<PY
import json
d=json.loads(open("$OUT/r03_prio$p.json").read().strip().splitlines()[-1])
print("prio=$p", d["value"], d["ms_per_step"], json.dumps(d["phases_ms_per_sweep"]))
PY
done
