#!/bin/bash
# Round-2 sixth visit (1 GPU): AIS with the series form of the increment + stream-ordered descriptor uploads.
TAG=${1:-r02_f}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest -m gpu -q --timeout=300 --durations=5 "tests/test_zz_dbm_tc_gpu.py::test_ais_fine_ladder_takes_the_series_form_and_matches_the_emulation" \
  "tests/test_zz_dbm_tc_gpu.py::test_ais_matches_exact_enumeration" "tests/test_zz_dbm_tc_gpu.py::test_ais_at_the_benchmark_shape_is_within_one_nat_of_the_float64_oracle" \
  tests/test_rbm_gpu.py > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -12 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --config cfg4-ais > $OUT/${TAG}_bench_cfg4-ais.json 2> $OUT/${TAG}_bench_cfg4-ais.err
python - <<PY
import json
d=json.load(open('$OUT/${TAG}_bench_cfg4-ais.json'))
print('  ', d['metric'], '%.4g'%d['value'], 'ms/ladder %.2f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'roofline %.3f (%.0f TF/s), step_frac %.3f'%(d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['step_frac']), d['quality'], d['clocks'])
PY
BM_TC_PROGRAM_TIMELINE=2 timeout 120 python tools/bench_configs.py cfg4-ais --dbm-compute bf16 --ais-runs 20000 --ais-betas 1000 > $OUT/${TAG}_ais_timeline.txt 2>&1
head -c 600 $OUT/${TAG}_ais_timeline.txt | tail -c 300; grep config $OUT/${TAG}_ais_timeline.txt
