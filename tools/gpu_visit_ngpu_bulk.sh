#!/bin/bash
# N-GPU visit: the peer-memory exchange with bulk asynchronous copies -- parity (dist_check), then the cfg2 bench line with the
# exchange's event profile.      usage: tools/gpu_visit_ngpu_bulk.sh <tag> <n>
TAG=${1:-r02_j}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
parse() { python - "$1" <<'PY'
import json, sys
s = open(sys.argv[1]).read()
i = s.find('{"metric"')
if i < 0:
    print('   no JSON line in', sys.argv[1]); sys.exit(0)
d = json.loads(s[i:].splitlines()[0])
print('   N=%d ms/step %.4f value %.4g launches/step %.1f e2e %.4g (%s steps) epoch-call %.4g' % (d['n_gpus'], d['ms_per_step'], d['value'], d['gpu_launches'] / d['steps'], d['e2e']['value'], d['e2e'].get('steps'), d.get('e2e_epoch_call', {}).get('value', float('nan'))))
PY
}
timeout 200 bash -c "$(declare -f run); N=$N; run 29675 tools/dist_check.py" > $OUT/${TAG}_dist_check_bulk.log 2>&1
echo "dist_check (bulk copies, $N ranks) exit $?"; grep -E "^rank 0|Error|Traceback" $OUT/${TAG}_dist_check_bulk.log | head -6
BM_PEER_PROFILE=1 timeout 200 bash -c "$(declare -f run); N=$N; run 29671 bench.py --gpus $N --steps 300 --warmup 10" > $OUT/${TAG}_bench_n${N}_bulk.json 2> $OUT/${TAG}_bench_n${N}_bulk.err
echo "bench N=$N bulk exit $?"; parse $OUT/${TAG}_bench_n${N}_bulk.json; grep "bm peer" $OUT/${TAG}_bench_n${N}_bulk.err | head -3
