#!/bin/bash
# Last visit of the round (1 GPU, short): smoke(), the fast parity files, the headline bench line with its default arguments.
TAG=${1:-r02_m}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
timeout 400 python -m pytest tests/test_rbm_gpu.py tests/test_tc_gpu.py tests/test_cabi.py tests/test_plugin_gpu.py tests/test_rbm_host.py -m gpu -q --timeout=120 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -c 3500 $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_reference.json 2>&1; echo "reference arm exit $?"; tail -c 900 $OUT/${TAG}_bench_reference.json
