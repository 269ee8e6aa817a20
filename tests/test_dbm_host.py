"""DBM host logic (composition from RBMs, fit loop, session semantics of the read-only queries,
save/load) on the numpy oracle (CPU) and on the CUDA engine (`-m gpu`).  The reference has no DBM
tests; these follow the usage in examples/dbm_mnist.py:100-170."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from boltzmann_machines import DBM
from boltzmann_machines.rbm import BernoulliRBM
from boltzmann_machines.utils import RNG

V, H1, H2 = 20, 12, 8


def data(n=32):
    return (RNG(seed=1337).rand(n, V) < 0.3).astype(np.float32)


def make_rbms(X, tag=''):
    rbm1 = BernoulliRBM(n_visible=V, n_hidden=H1, max_epoch=2, batch_size=8, verbose=False, random_seed=1,
                        dbm_first=True, model_path='rbm1{0}/'.format(tag)).fit(X)
    Q = rbm1.transform(X)
    rbm2 = BernoulliRBM(n_visible=H1, n_hidden=H2, max_epoch=2, batch_size=8, verbose=False, random_seed=2,
                        dbm_last=True, model_path='rbm2{0}/'.format(tag)).fit(Q)
    return [rbm1, rbm2]


def make_dbm(rbms, path='dbm/', **kw):
    cfg = dict(n_particles=8, batch_size=8, max_epoch=2, n_gibbs_steps=2, max_mf_updates=5, mf_tol=1e-5,
               learning_rate=0.01, l2=1e-4, max_norm=4., sparsity_cost=[0.01, 0.02], sparsity_target=0.2,
               random_seed=7, verbose=False, model_path=path)
    cfg.update(kw)
    return DBM(rbms=rbms, **cfg)


def test_composition_halves_doubled_evidence(engines, workdir):
    X = data()
    rbms = make_rbms(X)
    dbm = make_dbm(rbms).init()
    w = dbm.get_tf_params(scope='weights')
    r1, r2 = [r.get_tf_params(scope='weights') for r in rbms]
    assert set(w) == {'vb', 'W', 'W_1', 'hb', 'hb_1'}
    assert_allclose(w['W'], r1['W']); assert_allclose(w['W_1'], r2['W'])
    assert_allclose(w['vb'], r1['vb'])
    assert_allclose(w['hb'], 0.5 * r1['hb'] + 0.5 * r2['vb'], rtol=1e-6)       # dbm.py:285-291
    assert_allclose(w['hb_1'], r2['hb'])
    assert dbm.n_layers_ == 2 and dbm.n_hiddens_ == [H1, H2]


def test_fit_is_deterministic_and_resumable(engines, workdir):
    X = data()
    rbms = make_rbms(X)
    a = make_dbm(rbms, 'a/').fit(X, X[:8])
    b = make_dbm(rbms, 'b/').fit(X, X[:8])
    wa, wb = a.get_tf_params(scope='weights'), b.get_tf_params(scope='weights')
    for k in wa:
        assert_allclose(wa[k], wb[k])
    assert a.epoch_ == 2 and a.iter_ == 8
    # continue in memory == reload from disk and continue
    a.set_params(max_epoch=3).fit(X)
    b2 = DBM.load_model('b/')
    b2.load_rbms(rbms)
    b2.set_params(max_epoch=3).fit(X)
    wa, wb = a.get_tf_params(scope='weights'), b2.get_tf_params(scope='weights')
    for k in wa:
        assert_allclose(wa[k], wb[k], rtol=1e-5, atol=1e-6)


def test_queries_leave_no_trace(engines, workdir):
    X = data()
    dbm = make_dbm(make_rbms(X)).fit(X)
    before = dbm.get_tf_params()
    T = dbm.transform(X)
    R = dbm.reconstruct(X)
    S = dbm.sample_v(n_gibbs_steps=3)
    lp = dbm.log_proba(X, log_Z=0.)
    assert T.shape == (len(X), H2) and R.shape == X.shape and S.shape == (8, V) and lp.shape == (len(X),)
    assert (T >= 0).all() and (T <= 1).all() and (R >= 0).all() and (R <= 1).all()
    after = dbm.get_tf_params()
    for k in before:
        assert_allclose(before[k], after[k], err_msg=k)
    # sample_v(save_model=True) does advance the persistent chains
    dbm.sample_v(n_gibbs_steps=3, save_model=True)
    assert dbm.n_samples_generated_ == 3
    assert not np.allclose(before['negative_particles/v'], dbm.get_tf_params()['negative_particles/v'])


def test_log_Z_small_model_close_to_exact(engines, workdir):
    """AIS against brute-force enumeration on a model small enough to enumerate (V=6, H1=4, H2=3)."""
    rng = np.random.RandomState(0)
    v, h1, h2 = 6, 4, 3
    X = (rng.rand(16, v) < 0.4).astype(np.float32)
    r1 = BernoulliRBM(n_visible=v, n_hidden=h1, max_epoch=1, batch_size=8, verbose=False, random_seed=1,
                      W_init=0.5 * rng.randn(v, h1), dbm_first=True, model_path='s1/').init()
    r2 = BernoulliRBM(n_visible=h1, n_hidden=h2, max_epoch=1, batch_size=8, verbose=False, random_seed=2,
                      W_init=0.5 * rng.randn(h1, h2), dbm_last=True, model_path='s2/').init()
    dbm = make_dbm([r1, r2], 'sdbm/', n_particles=4).init()
    w = dbm.get_tf_params(scope='weights')
    W0, W1 = w['W'].astype(np.float64), w['W_1'].astype(np.float64)
    vb, c0, c1 = w['vb'].astype(np.float64), w['hb'].astype(np.float64), w['hb_1'].astype(np.float64)
    # exact: sum over h1 of exp(h1.c0) prod_v (1 + exp(W0 h1 + vb)) prod_h2 (1 + exp(h1 W1 + c1))
    terms = []
    for s in range(2 ** h1):
        x = np.array([(s >> i) & 1 for i in range(h1)], dtype=np.float64)
        terms.append(x @ c0 + np.logaddexp(0, W0 @ x + vb).sum() + np.logaddexp(0, x @ W1 + c1).sum())
    exact = np.logaddexp.reduce(terms)
    log_mean, (lo, hi), values = dbm.log_Z(n_betas=2000, n_runs=64, n_gibbs_steps=1)
    assert values.shape == (64,)
    assert abs(log_mean - exact) < 0.05, (log_mean, exact)
    assert lo <= log_mean <= hi


def test_ais_runs_shard_by_first_run():
    """AIS runs are independent chains -- run r draws from row r of the AIS sites, so any split of the
    ladder over ranks / calls reproduces the unsharded runs exactly (the engine's bm_dbm_ais / bm_dbm_ais_rows)."""
    from oracle.dbm import OracleDBM
    from oracle.dbm_bf16 import OracleDBMbf16
    for cls in (OracleDBM, OracleDBMbf16):
        cfg = dict(n_visible=7, n_hiddens=[5, 4], dtype='float32', n_particles=4, batch_size=4)
        ora = cls(cfg)
        rng = np.random.RandomState(0)
        ora.set_params({'W': 0.3 * rng.randn(7, 5), 'W_1': 0.3 * rng.randn(5, 4), 'vb': 0.1 * rng.randn(7),
                        'hb': 0.1 * rng.randn(5), 'hb_1': 0.1 * rng.randn(4)})
        full = ora.ais(13, 40, 2, 99)
        parts = np.concatenate([ora.ais(5, 40, 2, 99, first_run=0), ora.ais(1, 40, 2, 99, first_run=5),
                                ora.ais(7, 40, 2, 99, first_run=6)])
        np.testing.assert_allclose(full, parts, rtol=0, atol=1e-6)    # same chains; a different chain would differ by ~0.1
        assert ora.ais(13, 40, 2, 99).tolist() == full.tolist()          # first_run does not leak into later calls
