"""The data-parallel paths of the engines on the CPU, through the library's own host code: tools/dist_check.py under torchrun
with two ranks, the stand-in CUDA runtime with interpreted kernels (tests/hostsim) and the shared-memory stand-in for NCCL
(BM_NCCL_LIB).  Checked there: data-parallel CD-k of the RBM engines against the single-process oracle (fp32 and bf16), the AIS
ladder sharded over the ranks against the unsharded one, the data-parallel DBM step (rows and persistent particles sharded,
max-allreduce in the mean-field test, one sum-allreduce of the statistics) against the oracle holding the global batch, and
the tensor-core DBM engine sharded against itself in one piece."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_through_the_host_simulation():
    pytest.importorskip('torch')
    obj = os.path.join(ROOT, 'boltzmann-machines_b200', 'build')
    if not (os.path.isdir(obj) and any(f.endswith('.o') for f in os.listdir(obj))):
        pytest.skip('library objects not built (run build.sh / __graft_entry__.build())')
    subprocess.check_call(['bash', os.path.join(ROOT, 'tests', 'hostsim', 'build.sh')], stdout=subprocess.DEVNULL)
    env = dict(os.environ, BM_HOSTSIM='1', BM_NCCL_LIB=os.path.join(ROOT, 'tests', 'hostsim', '_build', 'libfakenccl.so'))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(29800 + os.getpid() % 100), os.path.join(ROOT, 'tools', 'dist_check.py')]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-4000:]
    import re
    for what in ('compute=fp32', 'compute=bf16', 'epoch compute=fp32', 'epoch compute=bf16', 'AIS:', 'DBM data parallel', 'tensor-core DBM data parallel'):
        n = len(re.findall(r'rank \d ' + re.escape(what), res.stdout))
        assert n == 2, (what, n, res.stdout[-2000:])
