#!/bin/bash
# Last visit of the round (1 GPU): smoke(), the headline bench line with its default arguments, the reference arm, then the whole
# GPU suite (4 pytest-xdist workers on the one GPU: the suite's wall time is the CPU oracle's, not the GPU's).
TAG=${1:-r02_n}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -c 3500 $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err
timeout 150 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_reference.json 2>&1; echo "reference arm exit $?"; tail -c 600 $OUT/${TAG}_bench_reference.json
timeout ${PYTEST_LIMIT:-420} python -m pytest tests -m gpu -q -n 4 --timeout=300 --durations=12 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -22 $OUT/${TAG}_pytest.log
