"""Run under torchrun (one process per GPU): data-parallel CD-k through the engine's NCCL allreduce
must reproduce the single-process oracle on the concatenated batch, on every rank."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'boltzmann-machines_b200'))


def main():
    import torch.distributed as dist
    from boltzmann_machines import _native
    from oracle.rbm import OracleRBM
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    if os.environ.get('BM_HOSTSIM') == '1':
        # dry run without GPUs (tests/hostsim): the library's host code on the stand-in runtime, kernels interpreted on the CPU,
        # collectives through the shared-memory stand-in for NCCL (BM_NCCL_LIB)
        import ctypes as C
        sim = _native.load_library(os.path.join(ROOT, 'tests', 'hostsim', '_build', 'libbm_hostsim.so'))
        sim.fakecuda_set_execute(1)
        _native._lib = sim
    dist.init_process_group('gloo')
    ctx = _native.Context(local)
    uid = [_native.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    rng = np.random.RandomState(0)
    V, H, rows = 784, 256, 128
    X = (rng.rand(rows * world, V) < 0.2).astype(np.float32)
    init = dict(W=(0.05 * rng.randn(V, H)).astype(np.float32), vb=(0.1 * rng.randn(V)).astype(np.float32),
                hb=np.zeros(H, np.float32))
    ok = True
    for compute, tol in (('fp32', 2e-5), ('bf16', 3e-3)):
        cfg = dict(n_visible=V, n_hidden=H, sample_v=False, sample_h=True, l2=1e-4, sparsity_cost=0.01, max_batch=rows,
                   compute=compute, dtype='float32')
        eng = _native.CudaRBM(cfg, ctx=ctx)
        ora = OracleRBM(cfg)
        eng.set_params(init); ora.set_params(init)
        for it in range(3):
            eng.train_step(X[rank * rows:(rank + 1) * rows], 0.05, 0.5, 2, 77, it)
            ora.train_step(X, 0.05, 0.5, 2, 77, it)          # the whole global batch in one process
        got, want = eng.get_params(['W', 'vb', 'hb']), ora.get_params(['W', 'vb', 'hb'])
        err = max(float(np.max(np.abs(got[k] - want[k]))) for k in got)
        # every rank must hold bit-identical parameters (same allreduced statistics, same update)
        mine = np.concatenate([got[k].ravel() for k in sorted(got)])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        identical = all(np.array_equal(gathered[0], g) for g in gathered)
        print('rank {0} compute={1}: max |engine - oracle(full batch)| = {2:.3e}, ranks identical: {3}'.format(
            rank, compute, err, identical), flush=True)
        ok = ok and err < tol and identical
        eng.close()
        # the whole-epoch entry point, data parallel: every rank feeds ITS rows of each global batch from a packed, page-locked
        # training set (bytes here: bm_rbm_train_epoch_u8; upload and conversion of batch i+1 overlap batch i); the oracle steps
        # through the concatenated batches
        eng, ora = _native.CudaRBM(cfg, ctx=ctx), OracleRBM(cfg)
        eng.set_params(init); ora.set_params(init)
        nb = 3
        Xe = (np.random.RandomState(5).rand(nb, world, rows, V) < 0.2).astype(np.float32)
        mine_rows = np.ascontiguousarray(Xe[:, rank].reshape(nb * rows, V))
        P = eng.pin(mine_rows)
        got_m = eng.train_epoch(P, rows, 0.05, 0.5, 2, 77, 100, metrics=('msre',), every=1)
        want_m = [ora.train_step(Xe[i].reshape(world * rows, V), 0.05, 0.5, 2, 77, 100 + i, metrics=('msre',))['msre'] for i in range(nb)]
        eng.unpin(P)
        got, want = eng.get_params(['W', 'vb', 'hb']), ora.get_params(['W', 'vb', 'hb'])
        err = max(float(np.max(np.abs(got[k] - want[k]))) for k in got)
        mine = np.concatenate([got[k].ravel() for k in sorted(got)])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        identical = all(np.array_equal(gathered[0], g) for g in gathered)
        print('rank {0} epoch compute={1}: max |engine - oracle(full batches)| = {2:.3e}, ranks identical: {3}, feed {4}'.format(
            rank, compute, err, identical, P.dtype), flush=True)
        ok = ok and err < tol and identical and len(got_m['msre']) == nb and np.all(np.isfinite(got_m['msre']))
        eng.close()
    # AIS: the runs shard over the ranks inside bm_dbm_ais (one sum-allreduce of n_runs doubles gathers them);
    # every rank must receive the ladder a single GPU computes (run r draws from row r whichever rank owns it)
    from oracle.dbm import OracleDBM
    dcfg = dict(n_visible=20, n_hiddens=[12, 8], dtype='float32', n_particles=4, batch_size=4)
    dparams = {'W': (0.3 * rng.randn(20, 12)).astype(np.float32), 'W_1': (0.3 * rng.randn(12, 8)).astype(np.float32),
               'vb': (0.1 * rng.randn(20)).astype(np.float32), 'hb': (0.1 * rng.randn(12)).astype(np.float32),
               'hb_1': (0.1 * rng.randn(8)).astype(np.float32)}
    dbm, dora = _native.CudaDBM(dict(dcfg, compute='fp32'), ctx=ctx), OracleDBM(dcfg)
    dbm.set_params(dparams); dora.set_params(dparams)
    n_runs = 13                                    # does not divide evenly
    sharded = dbm.ais(n_runs, 50, 1, 4321)         # collective: every rank calls it
    alone = dbm.ais(n_runs, 50, 1, 4321, first_run=0)      # the same ladder computed entirely on this GPU
    want = dora.ais(n_runs, 50, 1, 4321)
    e1, e2 = float(np.max(np.abs(sharded - alone))), float(np.max(np.abs(sharded - want)))
    print('rank {0} AIS: max |sharded - single GPU| = {1:.3e}, max |sharded - oracle| = {2:.3e}'.format(rank, e1, e2), flush=True)
    ok = ok and e1 < 1e-6 and e2 < 5e-3
    dbm.close()
    # DBM training step, data parallel: batch rows and persistent particles sharded over the ranks (global particle
    # index in the draws, max over ranks in the mean-field test, one sum-allreduce of the statistics) must reproduce
    # the single-process oracle with `world` times the batch and the particles
    Bd, Md = 8, 6
    def dcfg2(B, M):
        return dict(n_visible=20, n_hiddens=[12, 8], v_kind='bernoulli', h_kinds=['bernoulli'] * 2, dtype='float32',
                    n_particles=M, batch_size=B, max_mf_updates=6, mf_tol=1e-3, l2=1e-4, max_norm=1.2, sample_v=True,
                    sample_h=[True, True], sparsity_target=[0.2, 0.1], sparsity_cost=[0.01, 0.005], sparsity_damping=0.8)
    Xd = (rng.rand(3, Bd * world, 20) < 0.3).astype(np.float32)
    deng, dref = _native.CudaDBM(dict(dcfg2(Bd, Md), compute='fp32'), ctx=ctx), OracleDBM(dcfg2(Bd * world, Md * world))
    for e in (deng, dref):
        e.set_params(dparams)
        e.init_particles(4242)
    dok = True
    for it in range(3):
        g = deng.train_step(Xd[it, rank * Bd:(rank + 1) * Bd], 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        w = dref.train_step(Xd[it], 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        dok = dok and g['n_mf_updates'] == w['n_mf_updates'] and abs(g['msre'] - w['msre']) < 1e-4 * max(w['msre'], 1e-9)
    got, want = deng.get_params(), dref.get_params()
    derr = 0.0
    for k in want:
        ref = want[k]
        if k in ('v', 'h', 'h_1'):
            ref = ref[rank * Md:(rank + 1) * Md]
        elif k.startswith('mu') and not k.startswith('mu_means'):
            ref = ref[rank * Bd:(rank + 1) * Bd]
        derr = max(derr, float(np.max(np.abs(got[k] - ref))))
    print('rank {0} DBM data parallel: max |engine - oracle(global)| = {1:.3e}, metrics agree: {2}'.format(rank, derr, dok), flush=True)
    ok = ok and dok and derr < 5e-5
    deng.close()
    # the same contract for the tensor-core DBM engine (compute='bf16'): sharded over the ranks it must reproduce ITSELF
    # run in one piece (a second context without communicator holds the global batch and particles on this GPU)
    solo = _native.Context(local)
    teng = _native.CudaDBM(dict(dcfg2(Bd, Md), compute='bf16'), ctx=ctx)
    tref = _native.CudaDBM(dict(dcfg2(Bd * world, Md * world), compute='bf16'), ctx=solo)
    for e in (teng, tref):
        e.set_params(dparams)
        e.init_particles(4242)
    tok = True
    for it in range(3):
        g = teng.train_step(Xd[it, rank * Bd:(rank + 1) * Bd], 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        w = tref.train_step(Xd[it], 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        tok = tok and g['n_mf_updates'] == w['n_mf_updates'] and abs(g['msre'] - w['msre']) < 1e-3 * max(w['msre'], 1e-9)
    got, want = teng.get_params(), tref.get_params()
    terr = 0.0
    for k in want:
        ref = want[k]
        if k in ('v', 'h', 'h_1'):
            ref = ref[rank * Md:(rank + 1) * Md]
        elif k.startswith('mu') and not k.startswith('mu_means'):
            ref = ref[rank * Bd:(rank + 1) * Bd]
        terr = max(terr, float(np.max(np.abs(got[k] - ref))))
    print('rank {0} tensor-core DBM data parallel: max |sharded - one piece| = {1:.3e}, metrics agree: {2}'.format(rank, terr, tok), flush=True)
    ok = ok and tok and terr < 2e-3
    teng.close(); tref.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
