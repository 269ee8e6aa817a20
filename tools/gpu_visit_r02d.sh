#!/bin/bash
# Round-2 fourth visit (1 GPU): AIS inside the epilogues (parity tests, bench), DBM programs as default, bf16 feed, the
# north-star gates at full size; ncu --set full captures of the program kernel at cfg2, cfg3 and inside an AIS ladder.
TAG=${1:-r02_d}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_zz_dbm_tc_gpu.py tests/test_rbm_gpu.py tests/test_dbm_gpu.py -m gpu -q --timeout=300 --durations=8 \
  "tests/test_full_size_gpu.py::test_north_star_gates_at_cfg2_after_three_epochs" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -30 $OUT/${TAG}_pytest.log
for c in cfg4-ais cfg4 cfg5-pcd cfg2; do
  timeout 300 python bench.py --config $c > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err
  echo "bench $c exit $?"; python - <<PY
import json
try:
    d=json.load(open('$OUT/${TAG}_bench_$c.json'))
    print('  ', d['metric'], '%.4g'%d['value'], d['unit'], 'ms/step %.4f'%d['ms_per_step'], 'e2e %.4g (%s)'%(d['e2e']['value'], d['e2e'].get('ms_per_step')), 'roofline %.3f (%.0f TF/s), step_frac %.3f'%(d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['step_frac']), 'launches/step %.1f'%(d['gpu_launches']/d['steps']), d['quality'], {k: d[k]['value'] for k in d if k.startswith('e2e_')}, d['clocks'])
except Exception as e:
    print('   no JSON:', e)
PY
  tail -2 $OUT/${TAG}_bench_$c.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_program_kernel -s 6 -c 1 -f -o $OUT/${TAG}_cfg2_tc_program \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_cfg2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_program_kernel -s 6 -c 1 -f -o $OUT/${TAG}_cfg3_tc_program \
  python bench.py --config cfg3 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_cfg3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_program_kernel -s 40 -c 1 -f -o $OUT/${TAG}_ais_tc_program \
  python tools/bench_configs.py cfg4-ais --dbm-compute bf16 --ais-runs 20000 --ais-betas 200 > $OUT/${TAG}_ncu_ais.log 2>&1
ls -la $OUT | tail -6
