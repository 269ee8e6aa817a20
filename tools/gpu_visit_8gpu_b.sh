#!/bin/bash
# 8-GPU visit b: the exchange with only the bf16 shadow all-gathered (fp32 rows pulled on demand): parity, then the cfg2 line.
TAG=${1:-r02_l}
N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 150 bash -c "$(declare -f run); N=$N; run 29671 tools/dist_check.py" > $OUT/${TAG}_dist_check.log 2>&1
echo "dist_check ($N ranks) exit $?"; grep -E "^rank 0|Error|Traceback" $OUT/${TAG}_dist_check.log | head -6
BM_BENCH_FIT_STEPS=400 timeout 150 bash -c "$(declare -f run); N=$N; run 29673 bench.py --gpus $N --steps 300 --warmup 10" > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err
echo "bench N=$N exit $?"
python - "$OUT/${TAG}_bench_n${N}.json" <<'PY'
import json, sys
s = open(sys.argv[1]).read()
i = s.find('{"metric"')
if i < 0:
    print('   no JSON line'); sys.exit(0)
d = json.loads(s[i:].splitlines()[0])
print('   N=%d ms/step %.4f value %.4g launches/step %.1f e2e %.4g (%s steps) epoch-call %.4g' % (d['n_gpus'], d['ms_per_step'], d['value'], d['gpu_launches'] / d['steps'], d['e2e']['value'], d['e2e'].get('steps'), d.get('e2e_epoch_call', {}).get('value', float('nan'))))
PY
tail -2 $OUT/${TAG}_bench_n${N}.err
