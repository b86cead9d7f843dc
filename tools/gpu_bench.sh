#!/bin/bash
# bench only (short): headline + phases
TAG=${1:-r03x}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python bench.py --steps 5 --warmup 4 --two-chains 0 --bfield-series 0 --extra-configs 0 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], json.dumps(d["phases_ms_per_sweep"]), json.dumps(d["checks"]["max_propagation_error"]), d["checks"]["g_vs_oracle_rel"])
PY
