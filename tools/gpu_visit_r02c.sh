#!/bin/bash
# Round-2 third visit (1 GPU): whole GPU suite (constant columns moved behind the last store box), the bench line of
# every BASELINE configuration.
TAG=${1:-r02_c}
OUT=gpurun_out
mkdir -p $OUT
timeout 1100 python -m pytest tests -m gpu -q -x --timeout=300 --durations=12 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -22 $OUT/${TAG}_pytest.log
for c in cfg2 cfg3 cfg5 cfg5-pcd cfg4 cfg4-ais; do
  timeout 300 python bench.py --config $c > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err
  echo "bench $c exit $?"; python - <<PY
import json
try:
    d=json.load(open('$OUT/${TAG}_bench_$c.json'))
    print('  ', d['metric'], '%.4g'%d['value'], d['unit'], 'ms/step %.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'roofline %.3f (%.0f TF/s), step_frac %.3f'%(d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['step_frac']), 'launches/step %.1f'%(d['gpu_launches']/d['steps']), d['quality'], 'cpu', d.get('cpu_baseline',{}).get('value'), d['clocks'])
except Exception as e:
    print('   no JSON:', e)
PY
  tail -2 $OUT/${TAG}_bench_$c.err
done
