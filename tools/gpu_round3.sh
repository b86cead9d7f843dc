#!/bin/bash
# Round-3-style validation pass (run through gpurun): full GPU suite, smoke, bench, stabilization breakdown.
TAG=${1:-r03b}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -s --durations=10 -p no:cacheprovider --timeout=400 --timeout-method=thread > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 4 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
timeout 200 python tools/stab_breakdown.py 16 > $OUT/${TAG}_stab.log 2>&1
cat $OUT/${TAG}_stab.log; tail -3 $OUT/${TAG}_smoke.log; tail -8 $OUT/${TAG}_pytest.log; grep -a FAIL $OUT/${TAG}_pytest.log | head
