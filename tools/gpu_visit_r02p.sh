#!/bin/bash
# Short 1-GPU visit: epochs convert the uploaded batch on the copy stream (double-buffered operand) + one-launch MSRE.
# Parity files first, then the bench line with the conversion on the copy stream (default) and on the compute stream.
TAG=${1:-r02_p}
OUT=gpurun_out
mkdir -p $OUT
timeout 150 python -m pytest tests/test_rbm_gpu.py tests/test_tc_gpu.py tests/test_rbm_host.py tests/test_full_size_gpu.py tests/test_z_reference_golden.py -m gpu -q -n 4 --timeout=120 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
show() { python - "$1" <<'PY'
import json, sys
t = open(sys.argv[1]).read(); d = json.loads(t[t.index('{"metric'):].splitlines()[0])
print('   value %.4g ms/step %.4f | e2e fit %.4g (%.4f ms) | epoch call %.4g (%.4f ms) | bf16 feed %.4g (%.4f ms)' % (
    d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_epoch_call']['value'], d['e2e_epoch_call']['ms_per_step'],
    d['e2e_float32']['value'], d['e2e_float32']['ms_per_step']))
PY
}
timeout 150 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_copy.json 2> $OUT/${TAG}_bench_copy.err; echo "bench (copy stream) exit $?"; show $OUT/${TAG}_bench_copy.json
BM_EPOCH_CONVERT_ON_COPY_STREAM=0 timeout 150 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_compute.json 2> $OUT/${TAG}_bench_compute.err; echo "bench (compute stream) exit $?"; show $OUT/${TAG}_bench_compute.json
