#!/bin/bash
# Scratch GPU visit used during kernel work: a test subset, an environment sweep of bench.py and one timeline.
#   gpurun --timeout 900 -- 'bash tools/probe_round.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_rbm_gpu.py -m gpu -q -rf > gpurun_out/probe_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/probe_pytest.log; tail -5 gpurun_out/probe_pytest.log
bash tools/sweep.sh "BM_X=1" "BM_TC_DW_OVERLAP=0" > gpurun_out/probe_sweep.txt 2>&1; cat gpurun_out/probe_sweep.txt
BM_TC_PROGRAM_TIMELINE=1 timeout 120 python tools/program_timeline.py 2>&1 | cut -c1-300 | grep -v "first-unit" > gpurun_out/probe_timeline.txt
