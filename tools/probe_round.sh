mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_rbm_gpu.py "tests/test_full_size_gpu.py::test_full_size_resident_epoch_is_deterministic_and_matches_fed_batches" "tests/test_full_size_gpu.py::test_full_size_step_properties[cfg2-cd5]" -m gpu -q -rf > gpurun_out/p4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/p4_pytest.log; tail -12 gpurun_out/p4_pytest.log | cut -c1-300
bash tools/sweep.sh "BM_X=1" "BM_TC_FLAGS=11" "BM_TC_POLL_NS=32" "BM_TC_POLL_NS=16" "BM_TC_EPI_NS=32" "BM_TC_FLAGS=11 BM_TC_POLL_NS=32 BM_TC_EPI_NS=32" "BM_TC_FLAGS=11 BM_TC_POLL_NS=20 BM_TC_EPI_NS=20 BM_TC_BN_DOWN=192" > gpurun_out/p4_sweep.txt 2>&1
cat gpurun_out/p4_sweep.txt
