"""What the tensor-core DBM engine (compute='bf16') is expected to produce, checked on the CPU: the bf16-operand
emulation (oracle/dbm_bf16.py) against the pinned storage-precision oracle (oracle/dbm.py) and exact enumeration.

The bounds asserted here are BASELINE.json's: MSRE within a few percent, AIS log Z within +-1.0 nat."""
import numpy as np
import pytest

from oracle.dbm import OracleDBM
from oracle.dbm_bf16 import OracleDBMbf16, softplus_diff
from oracle.rbm import bf16_round


def make_cfg(V=30, Hs=(18, 11), **kw):
    cfg = dict(n_visible=V, n_hiddens=list(Hs), v_kind='bernoulli', h_kinds=['bernoulli'] * len(Hs),
               h_n_samples=[100.] * len(Hs), dtype='float32', n_particles=12, batch_size=10, max_mf_updates=6, mf_tol=1e-6,
               l2=1e-4, max_norm=3.0, sample_v=True, sample_h=[True] * len(Hs),
               sparsity_target=[0.2] * len(Hs), sparsity_cost=[0.01] * len(Hs), sparsity_damping=0.9)
    cfg.update(kw)
    return cfg


def init(cfg, engines, seed=0, scale=0.3):
    rng = np.random.RandomState(seed)
    sizes = [cfg['n_visible']] + cfg['n_hiddens']
    d = {'vb': (0.1 * rng.randn(sizes[0])).astype(np.float32)}
    for i in range(len(cfg['n_hiddens'])):
        s = '' if i == 0 else '_%d' % i
        d['W' + s] = (scale * rng.randn(sizes[i], sizes[i + 1])).astype(np.float32)
        d['hb' + s] = (0.1 * rng.randn(sizes[i + 1])).astype(np.float32)
    for e in engines:
        e.set_params(d)
        e.init_particles(4242)
    return d


def batch(cfg, rows, seed=1):
    return (np.random.RandomState(seed).rand(rows, cfg['n_visible']) < 0.3).astype(np.float32)


def test_softplus_difference_has_no_cancellation():
    rng = np.random.RandomState(0)
    z = np.concatenate([rng.randn(4000) * 8, [-120., -30., 0., 30., 120.]]).astype(np.float32)
    for a, b in [(0., 1e-3), (0.25, 0.251), (0.999, 1.0), (0.5, 0.5)]:
        a32, b32 = np.float32(a), np.float32(b)
        want = np.logaddexp(0, np.float64(b32) * z.astype(np.float64)) - np.logaddexp(0, np.float64(a32) * z.astype(np.float64))
        got = softplus_diff(a, b, z).astype(np.float64)
        # relative accuracy of the (small) difference itself: a float32 softplus pair would lose ~4 digits here
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-12)
        assert np.all(np.isfinite(got))


@pytest.mark.parametrize('Hs', [(18,), (18, 11), (18, 11, 7)])
def test_bf16_training_tracks_the_pinned_oracle(Hs):
    cfg = make_cfg(Hs=Hs)
    ref, emu = OracleDBM(cfg), OracleDBMbf16(cfg)
    init(cfg, (ref, emu))
    for it in range(4):
        X = batch(cfg, 10, seed=it)
        a = ref.train_step(X, 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        b = emu.train_step(X, 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        assert b['msre'] == pytest.approx(a['msre'], rel=0.05)
    pa, pb = ref.get_params(), emu.get_params()
    for k in pa:
        if k.startswith(('W', 'vb', 'hb')):
            np.testing.assert_allclose(pb[k], pa[k], atol=2e-2, err_msg=k)
    # everything the engine keeps as a bf16 operand is representable in bf16
    for k in ('v', 'h', 'mu'):
        np.testing.assert_array_equal(pb[k], bf16_round(pb[k]))
    # sampled particles are binary
    assert set(np.unique(pb['h'])) <= {0.0, 1.0}


def test_bf16_queries_track_the_pinned_oracle():
    cfg = make_cfg()
    ref, emu = OracleDBM(cfg), OracleDBMbf16(cfg)
    init(cfg, (ref, emu))
    X = batch(cfg, 10)
    ref.train_step(X, 0.05, 0.5, 1, 5, 0); emu.train_step(X, 0.05, 0.5, 1, 5, 0)
    Xq = batch(cfg, 7, seed=9)
    np.testing.assert_allclose(emu.transform(Xq), ref.transform(Xq), atol=2e-2)
    np.testing.assert_allclose(emu.reconstruct(Xq), ref.reconstruct(Xq), atol=2e-2)
    np.testing.assert_allclose(emu.log_proba(Xq), ref.log_proba(Xq), atol=0.5)
    v = emu.sample_v(3, 11, 4)
    assert v.shape == (cfg['n_particles'], cfg['n_visible']) and np.all((v >= 0) & (v <= 1))


@pytest.mark.parametrize('k', [1, 3])
def test_bf16_ais_matches_exact_enumeration_and_the_pinned_oracle(k):
    cfg = make_cfg(V=7, Hs=(5, 4), n_particles=4, batch_size=4)
    ref, emu = OracleDBM(cfg), OracleDBMbf16(cfg)
    init(cfg, (ref, emu))
    a = emu.ais(32, 500, k, 2222)
    b = ref.ais(32, 500, k, 2222)
    lm = lambda v: np.logaddexp.reduce(v) - np.log(len(v))
    assert abs(lm(a) - lm(b)) < 0.1
    # exact log Z of the model the bf16 engine actually evaluates (weights rounded to bf16)
    p = emu.get_params()
    W0, W1 = bf16_round(p['W']).astype(np.float64), bf16_round(p['W_1']).astype(np.float64)
    terms = []
    for s in range(2 ** 5):
        x = np.array([(s >> i) & 1 for i in range(5)], dtype=np.float64)
        terms.append(x @ p['hb'] + np.logaddexp(0, W0 @ x + p['vb']).sum() + np.logaddexp(0, x @ W1 + p['hb_1']).sum())
    exact = np.logaddexp.reduce(terms)
    assert abs(lm(a) - exact) < 0.1, (lm(a), exact)


def test_bf16_ais_is_within_one_nat_at_mnist_width():
    """BASELINE.json: AIS log Z within +-1.0 of the reference path (784-64-32, 48 runs x 150 betas)."""
    cfg = make_cfg(V=784, Hs=(64, 32), n_particles=4, batch_size=4)
    ref, emu = OracleDBM(cfg), OracleDBMbf16(cfg)
    init(cfg, (ref, emu), scale=0.05)
    a = emu.ais(48, 150, 1, 1)
    b = ref.ais(48, 150, 1, 1)
    lm = lambda v: np.logaddexp.reduce(v) - np.log(len(v))
    assert abs(lm(a) - lm(b)) < 1.0, (lm(a), lm(b))
