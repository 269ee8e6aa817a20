"""2-GPU data parallelism through the C-ABI (NCCL allreduce of the gradient statistics)."""
import os
import subprocess
import sys

import pytest

from boltzmann_machines import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpus_reproduce_the_single_process_oracle():
    if _native.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29671', os.path.join(ROOT, 'tools', 'dist_check.py')]
    env = dict(os.environ)
    if env.get('BM_HOSTSIM') == '1':          # dry run without GPUs (tests/hostsim): collectives through the stand-in for NCCL
        env.setdefault('BM_NCCL_LIB', os.path.join(ROOT, 'tests', 'hostsim', '_build', 'libfakenccl.so'))
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-4000:]
