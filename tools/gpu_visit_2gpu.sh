#!/bin/bash
# 2-GPU visit (gpurun --gpus 2): data-parallel parity on hardware (tools/dist_check.py under torchrun: ranks bit-identical,
# equal to the single-process oracle; sharded AIS; DBM data parallelism) with the peer-memory exchange and with the
# ncclAllReduce fallback, then the bench at N = 1, 2 with both.     usage: tools/gpu_visit_2gpu.sh [tag] [n]
TAG=${1:-r02_c}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 300 bash -c "$(declare -f run); N=$N; run 29671 tools/dist_check.py" > $OUT/${TAG}_dist_check_peer.log 2>&1
echo "dist_check (peer memory) exit $?" | tee -a $OUT/${TAG}_dist_check_peer.log; grep -E "^rank|Error|error" $OUT/${TAG}_dist_check_peer.log | tail -24
BM_PEER=0 timeout 300 bash -c "$(declare -f run); N=$N; run 29672 tools/dist_check.py" > $OUT/${TAG}_dist_check_nccl.log 2>&1
echo "dist_check (BM_PEER=0) exit $?" | tee -a $OUT/${TAG}_dist_check_nccl.log; grep -E "^rank|Error|error" $OUT/${TAG}_dist_check_nccl.log | tail -24
timeout 200 python bench.py --gpus 1 --steps 300 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench N=1 exit $?"; python -c "import json;d=json.load(open('$OUT/${TAG}_bench_n1.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'])"
timeout 300 bash -c "$(declare -f run); N=$N; run 29673 bench.py --gpus $N --steps 300 --warmup 10" > $OUT/${TAG}_bench_n${N}_peer.json 2> $OUT/${TAG}_bench_n${N}_peer.err
echo "bench N=$N (peer memory) exit $?"; python -c "import json;d=json.load(open('$OUT/${TAG}_bench_n${N}_peer.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"; tail -3 $OUT/${TAG}_bench_n${N}_peer.err
BM_PEER=0 timeout 300 bash -c "$(declare -f run); N=$N; run 29674 bench.py --gpus $N --steps 300 --warmup 10" > $OUT/${TAG}_bench_n${N}_nccl.json 2> $OUT/${TAG}_bench_n${N}_nccl.err
echo "bench N=$N (BM_PEER=0) exit $?"; python -c "import json;d=json.load(open('$OUT/${TAG}_bench_n${N}_nccl.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
