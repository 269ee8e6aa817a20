"""Data-parallel contract on CPU (gloo, world_size 2): chains sharded by rows, Philox keyed by the
global row index, one sum-allreduce of the gradient statistics -> the sharded run reproduces the
single-process run on the concatenated batch.  (The engine implements the same contract with NCCL;
tests/test_multi_gpu.py checks it on 2 GPUs.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, X, init, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle.rbm import OracleRBM
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    cfg = dict(n_visible=X.shape[1], n_hidden=init['W'].shape[1], sample_v=True, sample_h=True, dropout=0.9, l2=1e-4,
               sparsity_cost=0.01)
    ora = OracleRBM(cfg)
    ora.set_params(init)
    rows = X.shape[0] // world
    for it in range(3):
        ora.train_step(X[rank * rows:(rank + 1) * rows], 0.05, 0.5, 2, 77, it, shard=(rank, world, allreduce))
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), **ora.get_params())
    dist.destroy_process_group()


def test_two_rank_oracle_equals_single_process(tmp_path):
    torch = pytest.importorskip('torch')
    import torch.multiprocessing as mp
    from oracle.rbm import OracleRBM
    rng = np.random.RandomState(0)
    X = (rng.rand(16, 24) < 0.3).astype(np.float32)
    init = dict(W=(0.1 * rng.randn(24, 10)).astype(np.float32), vb=np.zeros(24, np.float32), hb=np.zeros(10, np.float32))
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, X, init, str(tmp_path)), nprocs=2, join=True)
    single = OracleRBM(dict(n_visible=24, n_hidden=10, sample_v=True, sample_h=True, dropout=0.9, l2=1e-4, sparsity_cost=0.01))
    single.set_params(init)
    for it in range(3):
        single.train_step(X, 0.05, 0.5, 2, 77, it)
    want = single.get_params()
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r))
        for k in want:
            np.testing.assert_allclose(got[k], want[k], atol=2e-6, err_msg='rank %d %s' % (r, k))
