#!/bin/bash
# Short 1-GPU visit: the co-residency check of the dataflow programs on the real device, then the fast parity files and the bench line.
TAG=${1:-r02_o}
OUT=gpurun_out
mkdir -p $OUT
BM_TC_DEBUG_RESIDENCY=1 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/${TAG}_smoke.log
timeout 200 python -m pytest tests -m gpu -q -n 4 --timeout=120 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
timeout 200 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; head -c 400 $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err
