"""DBM engine (CUDA, through the C-ABI) against the numpy oracle on the same seeded inputs:
mean-field, PCD particles, the whole training step, the read-only queries and AIS."""
import numpy as np
import pytest

from boltzmann_machines import _native
from oracle.dbm import OracleDBM

pytestmark = pytest.mark.gpu


def make_cfg(V=30, Hs=(18, 11), dtype='float32', **kw):
    cfg = dict(n_visible=V, n_hiddens=list(Hs), v_kind='bernoulli', h_kinds=['bernoulli'] * len(Hs),
               h_n_samples=[100.] * len(Hs), dtype=dtype, compute='fp32', n_particles=12, batch_size=10, max_mf_updates=6, mf_tol=1e-6,
               l2=1e-4, max_norm=3.0, sample_v=True, sample_h=[True] * len(Hs),
               sparsity_target=[0.2] * len(Hs), sparsity_cost=[0.01] * len(Hs), sparsity_damping=0.9)
    cfg.update(kw)
    return cfg


def make_pair(cfg, seed=0):
    rng = np.random.RandomState(seed)
    eng, ora = _native.CudaDBM(cfg), OracleDBM(cfg)
    dt = np.dtype(cfg['dtype'])
    sizes = [cfg['n_visible']] + cfg['n_hiddens']
    d = {'vb': (0.1 * rng.randn(sizes[0])).astype(dt)}
    for i in range(len(cfg['n_hiddens'])):
        s = '' if i == 0 else '_%d' % i
        d['W' + s] = (0.3 * rng.randn(sizes[i], sizes[i + 1])).astype(dt)
        d['hb' + s] = (0.1 * rng.randn(sizes[i + 1])).astype(dt)
    for e in (eng, ora):
        e.set_params(d)
        e.init_particles(4242)
    return eng, ora


def batch(cfg, rows, seed=1):
    return (np.random.RandomState(seed).rand(rows, cfg['n_visible']) < 0.3).astype(cfg['dtype'])


def same(eng, ora, names=None, atol=2e-5):
    g, w = eng.get_params(names), ora.get_params(names)
    for k in w:
        np.testing.assert_allclose(g[k], w[k], atol=atol, err_msg=k)


@pytest.mark.parametrize('Hs', [(18,), (18, 11), (18, 11, 7)])
@pytest.mark.parametrize('dtype', ['float32', 'float64'])
def test_particle_init_and_training_steps(Hs, dtype):
    cfg = make_cfg(Hs=Hs, dtype=dtype)
    eng, ora = make_pair(cfg)
    same(eng, ora, ['v', 'h'], atol=1e-6)
    for it in range(3):
        X = batch(cfg, 10, seed=it)
        got = eng.train_step(X, 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        want = ora.train_step(X, 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        assert got['n_mf_updates'] == want['n_mf_updates']
        assert got['msre'] == pytest.approx(want['msre'], rel=1e-4)
    same(eng, ora, atol=5e-5 if dtype == 'float32' else 1e-9)
    eng.close()


@pytest.mark.parametrize('dtype', ['float32', 'float64'])
@pytest.mark.parametrize('v_kind,h_kinds', [('bernoulli', ('bernoulli', 'multinomial')), ('gaussian', ('bernoulli', 'multinomial')),
                                            ('bernoulli', ('multinomial', 'bernoulli')), ('gaussian', ('bernoulli', 'bernoulli'))])
def test_non_bernoulli_layers_train_and_query_like_the_oracle(v_kind, h_kinds, dtype):
    """Multinomial hidden layers and Gaussian visibles inside a DBM (examples/dbm_cifar_naive.py stacks a Gaussian-visible RBM
    and a multinomial top RBM; reference: dbm.py:385-427 with the layers of layers.py:54-92): training steps, then the queries."""
    cfg = make_cfg(V=24, Hs=(14, 9), dtype=dtype, v_kind=v_kind, h_kinds=list(h_kinds), h_n_samples=[100., 6.],
                   sparsity_cost=[0., 0.])
    if v_kind == 'gaussian':
        cfg['sigma'] = list(np.linspace(0.6, 1.4, 24))
    if 'multinomial' in h_kinds:
        cfg['h_n_samples'] = [6. if k == 'multinomial' else 100. for k in h_kinds]
    eng, ora = make_pair(cfg)
    tol = 1e-4 if dtype == 'float32' else (5e-6 if v_kind == 'gaussian' else 1e-9)    # float32 Box-Muller noise in float64 models
    same(eng, ora, ['v', 'h', 'h_1'], atol=1e-6 if v_kind == 'bernoulli' else tol)
    rng = np.random.RandomState(3)
    for it in range(3):
        X = rng.randn(10, 24).astype(dtype) if v_kind == 'gaussian' else batch(cfg, 10, seed=it)
        got = eng.train_step(X, 0.02, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        want = ora.train_step(X, 0.02, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        assert abs(got['n_mf_updates'] - want['n_mf_updates']) <= (1 if dtype == 'float32' else 0)
        assert got['msre'] == pytest.approx(want['msre'], rel=1e-3)
    same(eng, ora, atol=tol)
    Xq = rng.randn(7, 24).astype(dtype) if v_kind == 'gaussian' else batch(cfg, 7, seed=9)
    np.testing.assert_allclose(eng.transform(Xq), ora.transform(Xq), atol=tol)
    np.testing.assert_allclose(eng.reconstruct(Xq), ora.reconstruct(Xq), atol=tol)
    np.testing.assert_allclose(eng.sample_v(2, 11, 4), ora.sample_v(2, 11, 4), atol=tol)
    eng.close()


def test_queries_match_oracle():
    cfg = make_cfg()
    eng, ora = make_pair(cfg)
    X = batch(cfg, 10)
    eng.train_step(X, 0.05, 0.5, 1, 5, 0); ora.train_step(X, 0.05, 0.5, 1, 5, 0)
    Xq = batch(cfg, 7, seed=9)
    np.testing.assert_allclose(eng.transform(Xq), ora.transform(Xq), atol=2e-5)
    np.testing.assert_allclose(eng.reconstruct(Xq), ora.reconstruct(Xq), atol=2e-5)
    np.testing.assert_allclose(eng.log_proba(Xq), ora.log_proba(Xq), rtol=1e-5, atol=1e-4)
    g, w = eng.val_metrics(Xq, 2, 7, 3), ora.val_metrics(Xq, 2, 7, 3)
    assert g['n_mf_updates'] == w['n_mf_updates'] and g['msre'] == pytest.approx(w['msre'], rel=1e-4)
    np.testing.assert_allclose(eng.sample_v(3, 11, 4), ora.sample_v(3, 11, 4), atol=2e-5)
    same(eng, ora, ['v', 'h', 'h_1', 'mu', 'mu_1'])
    eng.close()


def test_ais_matches_oracle_and_exact_enumeration():
    cfg = make_cfg(V=7, Hs=(5, 4), n_particles=4, batch_size=4)
    eng, ora = make_pair(cfg)
    a = eng.ais(32, 500, 1, 2222)
    b = ora.ais(32, 500, 1, 2222)
    # same Philox stream: the runs agree individually up to float accumulation
    np.testing.assert_allclose(a, b, atol=5e-3)
    p = ora.get_params()
    W0, W1 = p['W'].astype(np.float64), p['W_1'].astype(np.float64)
    terms = []
    for s in range(2 ** 5):
        x = np.array([(s >> i) & 1 for i in range(5)], dtype=np.float64)
        terms.append(x @ p['hb'] + np.logaddexp(0, W0 @ x + p['vb']).sum() + np.logaddexp(0, x @ W1 + p['hb_1']).sum())
    exact = np.logaddexp.reduce(terms)
    est = np.logaddexp.reduce(a) - np.log(len(a))
    assert abs(est - exact) < 0.1, (est, exact)
    eng.close()


def test_ais_paper_scale_is_within_one_nat_of_the_oracle():
    """BASELINE.json: AIS log Z within +-1.0 of the reference path (here at 784-64-32, 64 runs x 200 betas)."""
    cfg = make_cfg(V=784, Hs=(64, 32), n_particles=4, batch_size=4)
    eng, ora = make_pair(cfg)
    a = eng.ais(64, 200, 1, 1)
    b = ora.ais(64, 200, 1, 1)
    lm = lambda v: np.logaddexp.reduce(v) - np.log(len(v))
    assert abs(lm(a) - lm(b)) < 1.0
    eng.close()


@pytest.mark.parametrize('dtype', ['float32', 'float64'])
def test_ais_runs_shard_by_first_run(dtype):
    """bm_dbm_ais_rows: runs [first, first + n) of the ladder, computed alone, equal those runs of the whole ladder
    (what lets bm_dbm_ais shard n_runs over the ranks of a communicator) -- and the oracle's."""
    cfg = make_cfg(V=7, Hs=(5, 4), n_particles=4, batch_size=4, dtype=dtype)
    eng, ora = make_pair(cfg)
    full = eng.ais(13, 40, 2, 99)
    parts = np.concatenate([eng.ais(5, 40, 2, 99, first_run=0), eng.ais(1, 40, 2, 99, first_run=5),
                            eng.ais(7, 40, 2, 99, first_run=6)])
    np.testing.assert_allclose(full, parts, rtol=0, atol=1e-6)    # same chains; a different chain would differ by ~0.1
    np.testing.assert_allclose(eng.ais(7, 40, 2, 99, first_run=6), ora.ais(7, 40, 2, 99, first_run=6), atol=5e-3)
    eng.close()
