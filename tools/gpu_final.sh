#!/bin/bash
# Final validation pass (run through gpurun): wrap timing, full GPU suite, smoke, bench.  Everything under a timeout.
TAG=${1:-r02z}
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python tools/wrap_bench.py 16 > $OUT/${TAG}_wrap.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --durations=10 -p no:cacheprovider --timeout=400 --timeout-method=thread > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 4 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
cat $OUT/${TAG}_wrap.log; tail -3 $OUT/${TAG}_smoke.log; tail -5 $OUT/${TAG}_pytest.log
