#!/bin/bash
# 8-GPU visit (gpurun --gpus 8): the data-parallel step over peer memory at the scale the driver's scaling run uses --
# parity (tools/dist_check.py under torchrun, 8 ranks), then the cfg2 bench line at N = 8 with the peer exchange and with ncclAllReduce.
TAG=${1:-r02_h}
N=${2:-8}
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
print('   N=%d ms/step %.4f value %.4g launches/step %.1f e2e %.4g (%s steps) epoch-call %.4g clocks %s' % (
    d['n_gpus'], d['ms_per_step'], d['value'], d['gpu_launches'] / d['steps'], d['e2e']['value'], d['e2e'].get('steps'),
    d.get('e2e_epoch_call', {}).get('value', float('nan')), d['clocks']))
PY
}
timeout 300 bash -c "$(declare -f run); N=$N; run 29671 tools/dist_check.py" > $OUT/${TAG}_dist_check_peer.log 2>&1
echo "dist_check (peer memory, $N ranks) exit $?" | tee -a $OUT/${TAG}_dist_check_peer.log; grep -E "^rank 0|Error|error|Traceback" $OUT/${TAG}_dist_check_peer.log | tail -12
timeout 300 bash -c "$(declare -f run); N=$N; run 29673 bench.py --gpus $N --steps 300 --warmup 10" > $OUT/${TAG}_bench_n${N}_peer.json 2> $OUT/${TAG}_bench_n${N}_peer.err
echo "bench N=$N (peer memory) exit $?"; parse $OUT/${TAG}_bench_n${N}_peer.json; tail -2 $OUT/${TAG}_bench_n${N}_peer.err
BM_PEER=0 BM_BENCH_FIT_STEPS=300 timeout 300 bash -c "$(declare -f run); N=$N; run 29674 bench.py --gpus $N --steps 300 --warmup 10" > $OUT/${TAG}_bench_n${N}_nccl.json 2> $OUT/${TAG}_bench_n${N}_nccl.err
echo "bench N=$N (BM_PEER=0) exit $?"; parse $OUT/${TAG}_bench_n${N}_nccl.json
