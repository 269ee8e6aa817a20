"""Timings of the other BASELINE.json configurations (bench.py itself measures configs[1] only: they are parity cases
there, tests/test_full_size_gpu.py).  One JSON object per configuration on stdout:

    python tools/bench_configs.py [cfg3] [cfg4] [cfg4-ais] [cfg5]

cfg3: GaussianRBM 3072-5000, batch 2048, CD-1            (bf16 tensor-core program)
cfg4: DBM 784-512-1024, batch = particles = 1024, 25 mean-field updates, 1 Gibbs step per PCD update
      (--dbm-compute fp32: CUDA-core engine, default; bf16: the opt-in tensor-core engine)
cfg4-ais: AIS on that DBM, 20000 runs (or --ais-runs) x 1000 betas
cfg5: the per-GPU shard of BernoulliRBM 784-4096, 4096 particles, 25 Gibbs steps per update (CD chain of the RBM engine)
cfg5-dbm: the same shard as the reference would run PCD: its DBM class with one hidden layer (--dbm-compute, BM_DBM_PCD_PROGRAM)
FLOP counts follow SURVEY.md section 8(d).  Times: CUDA events on the engine's stream after warm-up.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'boltzmann-machines_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

from boltzmann_machines import _native       # noqa: E402


def timed(ctx, fn, steps, warmup):
    for i in range(warmup):
        fn(i)
    ctx.sync()
    ctx.timer_start()
    for i in range(steps):
        fn(warmup + i)
    ms = ctx.timer_stop()
    ctx.sync()
    return ms / steps


def rbm_case(name, kind, V, H, B, k, steps, warmup, lr):
    ctx = _native.Context.default()
    rng = np.random.RandomState(1)
    cfg = dict(n_visible=V, n_hidden=H, v_kind=kind, h_kind='bernoulli', dtype='float32', compute='bf16', l2=1e-5,
               sample_v=False, sample_h=True, max_batch=B)
    if kind == 'gaussian':
        cfg['sigma'] = np.ones(V)
        X = rng.randn(4 * B, V).astype(np.float32)
    else:
        X = (rng.rand(4 * B, V) < 0.13).astype(np.float32)
    eng = _native.CudaRBM(cfg)
    eng.init_normal_W(0.01 if kind != 'gaussian' else 0.0008, 1337)
    eng.set_data(X)
    ms = timed(ctx, lambda i: eng.train_step_at((i % 4) * B, B, lr, 0.5, k, 7, i), steps, warmup)
    flop = 2.0 * B * V * H * (2 * k + 3)
    out = dict(config=name, ms_per_step=ms, gibbs_updates_per_s=B * k / (ms * 1e-3), tflops=flop / (ms * 1e-3) / 1e12,
               flop_per_step=flop, steps=steps)
    eng.close()
    return out


def dbm(ctx, compute='fp32'):
    V, Hs, B = 784, [512, 1024], 1024
    cfg = dict(compute=compute, n_visible=V, n_hiddens=Hs, v_kind='bernoulli', h_kinds=['bernoulli'] * 2, h_n_samples=[100.] * 2,
               dtype='float32', n_particles=B, batch_size=B, max_mf_updates=25, mf_tol=1e-7, l2=1e-7, max_norm=6.0,
               sample_v=True, sample_h=[True, True], sparsity_target=[0.2, 0.1], sparsity_cost=[1e-4, 5e-5],
               sparsity_damping=0.9)
    rng = np.random.RandomState(2)
    eng = _native.CudaDBM(cfg)
    eng.set_params({'vb': np.zeros(V, np.float32), 'W': (0.02 * rng.randn(V, Hs[0])).astype(np.float32),
                    'hb': np.zeros(Hs[0], np.float32), 'W_1': (0.02 * rng.randn(Hs[0], Hs[1])).astype(np.float32),
                    'hb_1': np.zeros(Hs[1], np.float32)})
    eng.init_particles(4242)
    return eng, cfg


def cfg4(steps, warmup, compute='fp32'):
    ctx = _native.Context.default()
    eng, cfg = dbm(ctx, compute)
    B, V, (H1, H2) = 1024, 784, cfg['n_hiddens']
    X = (np.random.RandomState(3).rand(B, V) < 0.13).astype(np.float32)
    n_mf = [0.0]

    def step(i):
        n_mf[0] = eng.train_step(X, 2e-3, 0.5, 1, 99, i, metrics=('msre', 'n_mf_updates'))['n_mf_updates']
    ms = timed(ctx, step, steps, warmup)
    # as written in the reference (X W_0 recomputed every mean-field update)
    mf = 2.0 * B * (V * H1 + 2 * H1 * H2) * 25 + 2.0 * B * (V * H1 + H1 * H2)
    pcd = 2.0 * B * (2 * V * H1 + 2 * H1 * H2)
    grads = 2.0 * B * 2 * (V * H1 + H1 * H2)
    out = dict(config='cfg4 DBM 784-512-1024 step', engine=eng.compute, ms_per_step=ms, n_mf_updates=n_mf[0],
               tflops=(mf + pcd + grads) / (ms * 1e-3) / 1e12, flop_per_step=mf + pcd + grads, steps=steps)
    eng.close()
    return out


def cfg4_ais(n_runs, n_betas, compute='fp32'):
    ctx = _native.Context.default()
    eng, cfg = dbm(ctx, compute)
    V, (H1, H2) = 784, cfg['n_hiddens']
    eng.ais(256, 20, 1, 1)          # warm-up
    ctx.sync()
    ctx.timer_start()
    vals = eng.ais(n_runs, n_betas, 1, 2222)
    ms = ctx.timer_stop()
    flop = float(n_runs) * (n_betas - 1) * (4.0 * H1 * V + 4.0 * H1 * H2)      # fused count, SURVEY 8(d)
    lm = float(np.logaddexp.reduce(vals) - np.log(len(vals)))
    out = dict(config='cfg4 AIS {0} runs x {1} betas'.format(n_runs, n_betas), engine=eng.compute, ms=ms, tflops=flop / (ms * 1e-3) / 1e12,
               chain_transitions_per_s=n_runs * (n_betas - 1) / (ms * 1e-3), log_Z=lm)
    eng.close()
    return out


def cfg5_dbm(steps, warmup, compute):
    """BASELINE.json configs[4] as the reference would run it: PCD-25 on a single RBM = its DBM class with ONE hidden layer
    (README.md:96), 4096 persistent particles and 4096 batch rows per GPU."""
    ctx = _native.Context.default()
    V, H, B, k = 784, 4096, 4096, 25
    cfg = dict(compute=compute, n_visible=V, n_hiddens=[H], v_kind='bernoulli', h_kinds=['bernoulli'], h_n_samples=[100.],
               dtype='float32', n_particles=B, batch_size=B, max_mf_updates=1, mf_tol=1e-7, l2=1e-5, max_norm=1e9,
               sample_v=False, sample_h=[True], sparsity_target=[0.1], sparsity_cost=[0.], sparsity_damping=0.9)
    rng = np.random.RandomState(4)
    eng = _native.CudaDBM(cfg)
    eng.set_params({'vb': np.zeros(V, np.float32), 'W': (0.01 * rng.randn(V, H)).astype(np.float32), 'hb': np.zeros(H, np.float32)})
    eng.init_particles(4242)
    X = (rng.rand(B, V) < 0.13).astype(np.float32)
    ms = timed(ctx, lambda i: eng.train_step(X, 0.01, 0.5, k, 99, i), steps, warmup)
    flop = 2.0 * B * V * H * (2 * k + 3 + 1)          # k sweeps, positive pass + its init pass, two gradient GEMMs
    out = dict(config='cfg5 as a 1-layer DBM (PCD-25) 784-4096, 4096 particles', engine=eng.compute, ms_per_step=ms,
               gibbs_updates_per_s=B * k / (ms * 1e-3), tflops=flop / (ms * 1e-3) / 1e12, steps=steps,
               env={k_: os.environ.get(k_) for k_ in ('BM_DBM_PCD_PROGRAM', 'BM_DBM_MF_CHUNK')})
    eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('which', nargs='*', default=['cfg3', 'cfg4', 'cfg4-ais', 'cfg5'])
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--ais-runs', type=int, default=20000)
    ap.add_argument('--ais-betas', type=int, default=1000)
    ap.add_argument('--dbm-compute', default='fp32', choices=['fp32', 'bf16'],
                    help="DBM engine of cfg4 / cfg4-ais: fp32 CUDA cores (default) or the opt-in tensor-core engine")
    a = ap.parse_args()
    for w in a.which:
        if w == 'cfg3':
            r = rbm_case('cfg3 GaussianRBM 3072-5000 batch 2048 CD-1', 'gaussian', 3072, 5000, 2048, 1, a.steps, a.warmup, 5e-4)
        elif w == 'cfg5':
            r = rbm_case('cfg5 shard BernoulliRBM 784-4096 batch 4096 k=25', 'bernoulli', 784, 4096, 4096, 25, a.steps, a.warmup, 0.01)
        elif w == 'cfg4':
            r = cfg4(a.steps, a.warmup, a.dbm_compute)
        elif w == 'cfg5-dbm':
            r = cfg5_dbm(a.steps, a.warmup, a.dbm_compute)
        elif w == 'cfg4-ais':
            r = cfg4_ais(a.ais_runs, a.ais_betas, a.dbm_compute)
        else:
            raise SystemExit('unknown configuration ' + w)
        print(json.dumps(r), flush=True)


if __name__ == '__main__':
    main()
