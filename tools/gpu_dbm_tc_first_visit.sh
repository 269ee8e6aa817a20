#!/bin/bash
# First GPU visit of the opt-in tensor-core DBM engine (csrc/bm_dbm_tc.cuh): its gated parity tests under a short timeout
# (a deadlocked tcgen05 kernel must not hold the box), then cfg4 / cfg4-ais timings of both engines side by side.
# usage: tools/gpu_dbm_tc_first_visit.sh [tag]      -> gpurun_out/<tag>_dbm_tc_*.{log,json}
TAG=${1:-r02_a}
OUT=gpurun_out
mkdir -p $OUT
# the one kernel-level question first, alone and under a short timeout: a two-pair op whose pairs have different B layouts
BM_EXPERIMENTAL=1 timeout 180 python -m pytest tests/test_zz_dbm_tc_gpu.py -x -q -k raw_two_pair > $OUT/${TAG}_dbm_tc_mixed_b.log 2>&1
echo "mixed-layout op exit $?" >> $OUT/${TAG}_dbm_tc_mixed_b.log; tail -5 $OUT/${TAG}_dbm_tc_mixed_b.log
BM_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_dbm_tc_gpu.py -x -q > $OUT/${TAG}_dbm_tc_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_dbm_tc_pytest.log
tail -15 $OUT/${TAG}_dbm_tc_pytest.log
# random-shape fuzz of the verified engines (fp32 / bf16 RBM, fp32 DBM) against the oracles
BM_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_zz_engine_fuzz_gpu.py -q > $OUT/${TAG}_engine_fuzz_pytest.log 2>&1
echo "engine fuzz pytest exit $?" >> $OUT/${TAG}_engine_fuzz_pytest.log
tail -8 $OUT/${TAG}_engine_fuzz_pytest.log
for c in fp32 bf16; do
  timeout 600 python tools/bench_configs.py cfg4 cfg4-ais --dbm-compute $c --steps 20 --ais-runs 20000 --ais-betas 1000 \
    > $OUT/${TAG}_dbm_tc_bench_$c.json 2> $OUT/${TAG}_dbm_tc_bench_$c.err
  echo "bench_configs($c) exit $?"; cat $OUT/${TAG}_dbm_tc_bench_$c.json
done
BM_DBM_PCD_PROGRAM=1 BM_DBM_MF_CHUNK=5 timeout 600 python tools/bench_configs.py cfg4 cfg5-dbm --dbm-compute bf16 --steps 20 \
  > $OUT/${TAG}_dbm_tc_bench_programs.json 2> $OUT/${TAG}_dbm_tc_bench_programs.err
echo "bench_configs(programs) exit $?"; cat $OUT/${TAG}_dbm_tc_bench_programs.json
# micro-benchmarks for the cluster-resident chain: DSMEM push latency / throughput, and cta_group::2 pairs +
# commit multicast + a DSMEM-fed A tile inside one 8-CTA cluster
for u in dsmem_push cluster8_pair_mma; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/$u tools/ubench/$u.cu > $OUT/${TAG}_ubench_$u.log 2>&1 \
    && timeout 60 /tmp/$u >> $OUT/${TAG}_ubench_$u.log 2>&1
  echo "$u exit $?" >> $OUT/${TAG}_ubench_$u.log; cat $OUT/${TAG}_ubench_$u.log
done
