#!/bin/bash
# Closing 1-GPU visit of round 2: the whole GPU suite and the default bench line at the final state of the product code.
TAG=${1:-r02_q}
OUT=gpurun_out
mkdir -p $OUT
timeout 150 python -m pytest tests -m gpu -q -n 4 --timeout=120 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
timeout 150 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; head -c 300 $OUT/${TAG}_bench.json; echo; tail -2 $OUT/${TAG}_bench.err
