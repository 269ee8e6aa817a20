#!/bin/bash
# Round-2 second visit: the whole GPU suite on the fused update (statistics out of the dW GEMMs, one update kernel), with the
# tensor-core DBM engine's tests un-gated; the cfg2 bench line; launch lists of a cfg4 DBM step and of a short AIS ladder.
TAG=${1:-r02_b}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 --durations=25 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -40 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 300 --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 2500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_b.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file $OUT/${TAG}_ais_launches.csv \
  python tools/bench_configs.py cfg4-ais --dbm-compute bf16 --ais-runs 20000 --ais-betas 60 > $OUT/${TAG}_ncu_ais.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file $OUT/${TAG}_cfg4_launches.csv \
  python tools/bench_configs.py cfg4 --dbm-compute bf16 --steps 3 --warmup 3 > $OUT/${TAG}_ncu_cfg4.log 2>&1
BM_DBM_PCD_PROGRAM=1 BM_DBM_MF_CHUNK=5 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 120 --csv --log-file $OUT/${TAG}_cfg4p_launches.csv \
  python tools/bench_configs.py cfg4 --dbm-compute bf16 --steps 3 --warmup 3 > $OUT/${TAG}_ncu_cfg4p.log 2>&1
ls -la $OUT | tail -8
