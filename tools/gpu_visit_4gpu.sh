#!/bin/bash
# 4-GPU visit: where the peer-memory exchange spends its time (BM_PEER_PROFILE), all-gather from the shard's blocks vs the
# split variant; parity of the split variant.
TAG=${1:-r02_i}
N=${2:-4}
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
print('   N=%d ms/step %.4f value %.4g launches/step %.1f epoch-call %.4g' % (d['n_gpus'], d['ms_per_step'], d['value'], d['gpu_launches'] / d['steps'], d.get('e2e_epoch_call', {}).get('value', float('nan'))))
PY
}
for split in 0 1; do
  BM_PEER_SPLIT=$split BM_PEER_PROFILE=1 BM_BENCH_FIT_STEPS=40 timeout 200 bash -c "$(declare -f run); N=$N; run 2967$split bench.py --gpus $N --steps 300 --warmup 10" > $OUT/${TAG}_bench_n${N}_split$split.json 2> $OUT/${TAG}_bench_n${N}_split$split.err
  echo "bench N=$N BM_PEER_SPLIT=$split exit $?"; parse $OUT/${TAG}_bench_n${N}_split$split.json; grep "bm peer" $OUT/${TAG}_bench_n${N}_split$split.err | head -4
done
BM_PEER_SPLIT=1 timeout 200 bash -c "$(declare -f run); N=$N; run 29675 tools/dist_check.py" > $OUT/${TAG}_dist_check_split.log 2>&1
echo "dist_check (split all-gather, $N ranks) exit $?"; grep -E "^rank 0|Error|Traceback" $OUT/${TAG}_dist_check_split.log | head -6
