#!/bin/bash
# Round-2 fifth visit (1 GPU): where an AIS temperature step spends its time (timeline of CTA 0 for one program launch, ops per
# launch sweep, old path beside it, ncu --set full of one AIS program launch); tests written since the last visit.
TAG=${1:-r02_e}
OUT=gpurun_out
mkdir -p $OUT
for ops in 90 30 3; do
  BM_DBM_AIS_OPS=$ops timeout 120 python tools/bench_configs.py cfg4-ais --dbm-compute bf16 --ais-runs 20000 --ais-betas 200 > $OUT/${TAG}_ais_ops$ops.json 2>&1
  echo "ais ops/launch $ops:"; tail -1 $OUT/${TAG}_ais_ops$ops.json
done
timeout 120 python tools/bench_configs.py cfg4-ais --dbm-compute bf16 --ais-runs 2000 --ais-betas 200 > $OUT/${TAG}_ais_r2000.json 2>&1; echo "2000 runs:"; tail -1 $OUT/${TAG}_ais_r2000.json
BM_DBM_AIS_EPILOGUE=0 timeout 120 python tools/bench_configs.py cfg4-ais --dbm-compute bf16 --ais-runs 20000 --ais-betas 200 > $OUT/${TAG}_ais_old.json 2>&1; echo "old path:"; tail -1 $OUT/${TAG}_ais_old.json
BM_TC_PROGRAM_TIMELINE=3 timeout 120 python tools/bench_configs.py cfg4-ais --dbm-compute bf16 --ais-runs 20000 --ais-betas 40 > $OUT/${TAG}_ais_timeline.txt 2>&1
echo "timeline exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_program_kernel -s 2 -c 1 -f -o $OUT/${TAG}_ais_tc_program \
  python tools/bench_configs.py cfg4-ais --dbm-compute bf16 --ais-runs 20000 --ais-betas 100 > $OUT/${TAG}_ncu_ais.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/${TAG}_ncu_ais.log
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_plugin_gpu.py tests/test_rbm_gpu.py tests/test_zz_engine_fuzz_gpu.py -m gpu -q --timeout=300 --durations=5 \
  "tests/test_zz_dbm_tc_gpu.py::test_ais_at_the_benchmark_shape_is_within_one_nat_of_the_float64_oracle" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -15 $OUT/${TAG}_pytest.log
ls -la $OUT | tail -4
