#!/bin/bash
# Round-2 first GPU visit: the standard parity suite, then the tensor-core DBM engine's first run on a B200 (gated tests,
# each file in its own process under a timeout), the engine fuzz, timings of the other BASELINE configurations, the
# program timeline and the DSMEM micro-benchmarks.   usage: tools/gpu_visit_r02a.sh [tag]
TAG=${1:-r02_a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x --timeout=300 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
BM_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_zz_dbm_tc_gpu.py -q -k raw_two_pair --timeout=60 > $OUT/${TAG}_dbm_tc_mixed_b.log 2>&1
echo "mixed-layout op exit $?" >> $OUT/${TAG}_dbm_tc_mixed_b.log; tail -4 $OUT/${TAG}_dbm_tc_mixed_b.log
BM_EXPERIMENTAL=1 timeout 500 python -m pytest tests/test_zz_dbm_tc_gpu.py -q --timeout=120 > $OUT/${TAG}_dbm_tc_pytest.log 2>&1
echo "dbm tc pytest exit $?" >> $OUT/${TAG}_dbm_tc_pytest.log; tail -25 $OUT/${TAG}_dbm_tc_pytest.log
BM_EXPERIMENTAL=1 timeout 500 python -m pytest tests/test_zz_engine_fuzz_gpu.py -q --timeout=120 > $OUT/${TAG}_engine_fuzz_pytest.log 2>&1
echo "engine fuzz pytest exit $?" >> $OUT/${TAG}_engine_fuzz_pytest.log; tail -12 $OUT/${TAG}_engine_fuzz_pytest.log
timeout 200 python tools/bench_configs.py cfg3 cfg5 --steps 30 > $OUT/${TAG}_bench_cfg35.json 2> $OUT/${TAG}_bench_cfg35.err
echo "bench_configs(cfg3 cfg5) exit $?"; cat $OUT/${TAG}_bench_cfg35.json
timeout 200 python tools/bench_configs.py cfg4 cfg4-ais --dbm-compute fp32 --steps 20 --ais-runs 2000 --ais-betas 1000 \
  > $OUT/${TAG}_dbm_bench_fp32.json 2> $OUT/${TAG}_dbm_bench_fp32.err
echo "bench_configs(fp32) exit $?"; cat $OUT/${TAG}_dbm_bench_fp32.json
timeout 200 python tools/bench_configs.py cfg4 cfg4-ais --dbm-compute bf16 --steps 20 --ais-runs 2000 --ais-betas 1000 \
  > $OUT/${TAG}_dbm_bench_bf16.json 2> $OUT/${TAG}_dbm_bench_bf16.err
echo "bench_configs(bf16) exit $?"; cat $OUT/${TAG}_dbm_bench_bf16.json; tail -3 $OUT/${TAG}_dbm_bench_bf16.err
timeout 200 python tools/bench_configs.py cfg4-ais --dbm-compute bf16 --ais-runs 20000 --ais-betas 1000 \
  > $OUT/${TAG}_dbm_bench_bf16_ais20k.json 2> $OUT/${TAG}_dbm_bench_bf16_ais20k.err
echo "bench_configs(bf16 ais 20k) exit $?"; cat $OUT/${TAG}_dbm_bench_bf16_ais20k.json
BM_DBM_PCD_PROGRAM=1 BM_DBM_MF_CHUNK=5 timeout 200 python tools/bench_configs.py cfg4 cfg5-dbm --dbm-compute bf16 --steps 20 \
  > $OUT/${TAG}_dbm_bench_programs.json 2> $OUT/${TAG}_dbm_bench_programs.err
echo "bench_configs(programs) exit $?"; cat $OUT/${TAG}_dbm_bench_programs.json; tail -3 $OUT/${TAG}_dbm_bench_programs.err
BM_DBM_PCD_PROGRAM=1 BM_DBM_MF_CHUNK=25 timeout 200 python tools/bench_configs.py cfg4 --dbm-compute bf16 --steps 20 \
  > $OUT/${TAG}_dbm_bench_programs25.json 2> $OUT/${TAG}_dbm_bench_programs25.err
echo "bench_configs(programs, chunk 25) exit $?"; cat $OUT/${TAG}_dbm_bench_programs25.json
BM_TC_PROGRAM_TIMELINE=2 timeout 120 python tools/program_timeline.py > $OUT/${TAG}_timeline.txt 2>&1
echo "timeline exit $?"
for u in dsmem_push cluster8_pair_mma; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/$u tools/ubench/$u.cu > $OUT/${TAG}_ubench_$u.log 2>&1 \
    && timeout 60 /tmp/$u >> $OUT/${TAG}_ubench_$u.log 2>&1
  echo "$u exit $?" >> $OUT/${TAG}_ubench_$u.log; cat $OUT/${TAG}_ubench_$u.log
done
timeout 300 python bench.py --steps 300 --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 2500 $OUT/${TAG}_bench.json
