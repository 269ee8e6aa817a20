"""TEST INFRASTRUCTURE: the float64 oracle's AIS log-weights at BASELINE.json configs[3]'s shape (DBM 784-512-1024, 1000
betas), first 128 chains of seed 7 -- minutes of numpy, so they are committed as a fixture
(tests/golden/ais_benchmark_shape_oracle.json) instead of being recomputed inside the GPU test.  The chains are independent
(Philox counters carry the chain index), so tests/test_oracle_fixtures.py re-derives a few of them on the CPU to pin the file.

    python tests/golden/make_ais_benchmark_oracle.py [n_runs]
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'boltzmann-machines_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

SHAPE = dict(V=784, Hs=(512, 1024), n_betas=1000, n_gibbs_steps=1, seed=7, particle_seed=4242, weight_seed=0, weight_scale=0.02)


def weights(shape=SHAPE):
    """the parameters tests/test_zz_dbm_tc_gpu.py::init draws (RandomState(0), scale 0.02), in float32"""
    rng = np.random.RandomState(shape['weight_seed'])
    sizes = [shape['V']] + list(shape['Hs'])
    d = {'vb': (0.1 * rng.randn(sizes[0])).astype(np.float32)}
    for i in range(len(shape['Hs'])):
        s = '' if i == 0 else '_%d' % i
        d['W' + s] = (shape['weight_scale'] * rng.randn(sizes[i], sizes[i + 1])).astype(np.float32)
        d['hb' + s] = (0.1 * rng.randn(sizes[i + 1])).astype(np.float32)
    return d


def oracle(shape=SHAPE):
    from oracle.dbm import OracleDBM
    Hs = list(shape['Hs'])
    cfg = dict(n_visible=shape['V'], n_hiddens=Hs, v_kind='bernoulli', h_kinds=['bernoulli'] * 2, h_n_samples=[100.] * 2,
               dtype='float64', compute='fp32', n_particles=4, batch_size=4, max_mf_updates=6, mf_tol=1e-6, l2=1e-4, max_norm=3.0,
               sample_v=True, sample_h=[True] * 2, sparsity_target=[0.2] * 2, sparsity_cost=[0.01] * 2, sparsity_damping=0.9)
    ref = OracleDBM(cfg)
    ref.set_params({k: v.astype(np.float64) for k, v in weights(shape).items()})
    ref.init_particles(shape['particle_seed'])
    return ref


def log_weights(n_runs, first_run=0, shape=SHAPE):
    ref = oracle(shape)
    try:
        return np.asarray(ref.ais(n_runs, shape['n_betas'], shape['n_gibbs_steps'], shape['seed'], first_run=first_run), dtype=np.float64)
    except TypeError:
        assert first_run == 0
        return np.asarray(ref.ais(n_runs, shape['n_betas'], shape['n_gibbs_steps'], shape['seed']), dtype=np.float64)


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    v = log_weights(n)
    out = dict(shape={k: (list(x) if isinstance(x, tuple) else x) for k, x in SHAPE.items()}, n_runs=n,
               log_weights=[float(x) for x in v],
               generated_by='tests/golden/make_ais_benchmark_oracle.py (oracle/dbm.py, float64)')
    with open(os.path.join(HERE, 'ais_benchmark_shape_oracle.json'), 'w') as fh:
        json.dump(out, fh)
    print('log mean exp', np.logaddexp.reduce(v) - np.log(n), 'mean', v.mean(), 'std', v.std())
