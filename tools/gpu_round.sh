#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list, one ncu --set full capture of the dominant
# kernel.  Everything lands in gpurun_out/<tag>_*.   usage: tools/gpu_round.sh <tag> [steps]
TAG=${1:-r01_x}
STEPS=${2:-500}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
timeout 300 python bench.py --steps $STEPS --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 2500 $OUT/${TAG}_bench.json
if [ -z "$SKIP_NCU" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_b.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_program_kernel -s 4 -c 1 \
    -f -o $OUT/${TAG}_tc_program python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
  ls -la $OUT | tail -12
fi
