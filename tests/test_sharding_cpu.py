"""Data-parallel contract on CPU (gloo, world_size 2): chains sharded by rows, Philox keyed by the
global row index, one sum-allreduce of the gradient statistics -> the sharded run reproduces the
single-process run on the concatenated batch.  (The engine implements the same contract with NCCL;
tests/test_multi_gpu.py checks it on 2 GPUs.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rbm_cfg(kind, V, H):
    cfg = dict(n_visible=V, n_hidden=H, sample_v=True, sample_h=True, dropout=0.9, l2=1e-4, sparsity_cost=0.01)
    if kind == 'gaussian':
        cfg.update(v_kind='gaussian', h_kind='bernoulli', sigma=np.linspace(0.7, 1.3, V))
    elif kind == 'multinomial':
        cfg.update(v_kind='bernoulli', h_kind='multinomial', h_n_samples=7)
    return cfg


def _worker(rank, world, port, X, init, out_dir, kind='bernoulli'):
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

    ora = OracleRBM(_rbm_cfg(kind, X.shape[1], init['W'].shape[1]))
    ora.set_params(init)
    rows = X.shape[0] // world
    for it in range(3):
        ora.train_step(X[rank * rows:(rank + 1) * rows], 0.05, 0.5, 2, 77, it, shard=(rank, world, allreduce))
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), **ora.get_params())
    dist.destroy_process_group()


@pytest.mark.parametrize('kind', ['bernoulli', 'gaussian', 'multinomial'])
def test_two_rank_oracle_equals_single_process(tmp_path, kind):
    torch = pytest.importorskip('torch')
    import torch.multiprocessing as mp
    from oracle.rbm import OracleRBM
    rng = np.random.RandomState(0)
    X = (rng.randn(16, 24) if kind == 'gaussian' else (rng.rand(16, 24) < 0.3)).astype(np.float32)
    init = dict(W=(0.1 * rng.randn(24, 10)).astype(np.float32), vb=np.zeros(24, np.float32), hb=np.zeros(10, np.float32))
    port = 29500 + (os.getpid() % 1000) + {'bernoulli': 0, 'gaussian': 11, 'multinomial': 23}[kind]
    mp.spawn(_worker, args=(2, port, X, init, str(tmp_path), kind), nprocs=2, join=True)
    single = OracleRBM(_rbm_cfg(kind, 24, 10))
    single.set_params(init)
    for it in range(3):
        single.train_step(X, 0.05, 0.5, 2, 77, it)
    want = single.get_params()
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r))
        for k in want:
            np.testing.assert_allclose(got[k], want[k], atol=2e-6, err_msg='rank %d %s' % (r, k))


# ---------------------------------------------------------------------------------------------------------
# DBM: batch rows (mean-field) and persistent particles sharded over the ranks (csrc/bm_dbm.cu)
# ---------------------------------------------------------------------------------------------------------
def _dbm_cfg(B, M, v_kind='bernoulli'):
    cfg = dict(n_visible=14, n_hiddens=[9, 6], v_kind=v_kind, h_kinds=['bernoulli'] * 2, dtype='float32', n_particles=M,
               batch_size=B, max_mf_updates=6, mf_tol=1e-3, l2=1e-4, max_norm=1.2, sample_v=True, sample_h=[True, True],
               sparsity_target=[0.2, 0.1], sparsity_cost=[0.01, 0.005], sparsity_damping=0.8)
    if v_kind == 'gaussian':
        cfg['sigma'] = np.linspace(0.8, 1.2, 14)
    return cfg


def _dbm_init(rng):
    return {'W': (0.3 * rng.randn(14, 9)).astype(np.float32), 'W_1': (0.3 * rng.randn(9, 6)).astype(np.float32),
            'vb': (0.1 * rng.randn(14)).astype(np.float32), 'hb': (0.1 * rng.randn(9)).astype(np.float32),
            'hb_1': (0.1 * rng.randn(6)).astype(np.float32)}


def _dbm_worker(rank, world, port, X, init, v_kind, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle.dbm import OracleDBM
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)

    def allsum(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    def allmax(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.numpy()

    B = X.shape[1] // world
    ora = OracleDBM(_dbm_cfg(B, 5, v_kind))
    ora.set_shard(rank, world, allsum, allmax)
    ora.set_params(init)
    ora.init_particles(4242)
    logs = []
    for it in range(X.shape[0]):
        logs.append(ora.train_step(X[it, rank * B:(rank + 1) * B], 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates')))
    out = ora.get_params()
    out['msre'] = np.array([l['msre'] for l in logs]); out['n_mf'] = np.array([l['n_mf_updates'] for l in logs])
    np.savez(os.path.join(out_dir, 'dbm_rank%d.npz' % rank), **out)
    dist.destroy_process_group()


@pytest.mark.parametrize('v_kind', ['bernoulli', 'gaussian'])
def test_two_rank_dbm_oracle_equals_single_process(tmp_path, v_kind):
    """Two shards of 4 batch rows + 5 particles each == one process with batch 8 and 10 particles: same draws (global
    particle index), same number of mean-field sweeps (max over shards), same update (summed statistics)."""
    pytest.importorskip('torch')
    import torch.multiprocessing as mp
    from oracle.dbm import OracleDBM
    rng = np.random.RandomState(3)
    steps, B, world = 3, 4, 2
    X = (rng.randn(steps, B * world, 14) if v_kind == 'gaussian' else (rng.rand(steps, B * world, 14) < 0.3)).astype(np.float32)
    init = _dbm_init(rng)
    port = 29600 + (os.getpid() % 1000) + (7 if v_kind == 'gaussian' else 0)
    mp.spawn(_dbm_worker, args=(world, port, X, init, v_kind, str(tmp_path)), nprocs=world, join=True)
    single = OracleDBM(_dbm_cfg(B * world, 5 * world, v_kind))
    single.set_params(init)
    single.init_particles(4242)
    logs = [single.train_step(X[it], 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates')) for it in range(steps)]
    want = single.get_params()
    ranks = [np.load(os.path.join(str(tmp_path), 'dbm_rank%d.npz' % r)) for r in range(world)]
    for r, got in enumerate(ranks):
        np.testing.assert_array_equal(got['n_mf'], [l['n_mf_updates'] for l in logs])
        np.testing.assert_allclose(got['msre'], [l['msre'] for l in logs], rtol=1e-5)
        for k in want:
            if k in ('v', 'h', 'h_1'):          # this rank's slice of the global particles
                np.testing.assert_allclose(got[k], want[k][r * 5:(r + 1) * 5], atol=2e-6, err_msg='rank %d %s' % (r, k))
            elif k.startswith('mu') and not k.startswith('mu_means'):
                np.testing.assert_allclose(got[k], want[k][r * B:(r + 1) * B], atol=2e-6, err_msg='rank %d %s' % (r, k))
            else:
                np.testing.assert_allclose(got[k], want[k], atol=2e-6, err_msg='rank %d %s' % (r, k))
    for k in ('W', 'W_1', 'vb', 'hb', 'hb_1'):      # the ranks hold identical parameters
        np.testing.assert_array_equal(ranks[0][k], ranks[1][k])
