mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rbm_gpu.py tests/test_tc_gpu.py -m gpu -x -q > gpurun_out/p2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/p2_pytest.log; tail -15 gpurun_out/p2_pytest.log
bash tools/sweep.sh "BM_TC_FLAGS=1" "BM_TC_FLAGS=3" "BM_TC_FLAGS=5" > gpurun_out/p2_sweep.txt 2>&1
cat gpurun_out/p2_sweep.txt
BM_TC_FLAGS=5 BM_TC_PROGRAM_TIMELINE=1 python tools/program_timeline.py 2>&1 | cut -c1-400 | grep -v "first-unit" > gpurun_out/p2_timeline.txt
timeout 300 python bench.py --steps 500 --warmup 10 > gpurun_out/p2_bench.json 2> gpurun_out/p2_bench.err; tail -c 2600 gpurun_out/p2_bench.json; tail -5 gpurun_out/p2_bench.err
