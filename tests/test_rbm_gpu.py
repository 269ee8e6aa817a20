"""Parity of the CUDA engine (through the C-ABI) against the CPU oracle on the same seeded
inputs.  fp32/fp64 compute: element-wise to ~1e-5 (float: summation order and libm differ
by ulps; a Bernoulli draw can only differ where |u - p| is at rounding level).  bf16 compute:
against the oracle that rounds at the same points (operands in bf16, fp32 accumulate).
"""
import numpy as np
import pytest

from boltzmann_machines import _native
from oracle.rbm import OracleRBM

pytestmark = pytest.mark.gpu

ACTS = ('X', 'h0_means', 'h0_states', 'v_means', 'v_states', 'h_means')


def make_cfg(kind, V, H, B, dtype='float32', compute='fp32', **kw):
    cfg = dict(n_visible=V, n_hidden=H, dtype=dtype, compute=compute, l2=1e-4, max_batch=B,
               sample_v=True, sample_h=True, sparsity_cost=0.01, sparsity_target=0.2)
    if kind == 'bernoulli':
        cfg.update(v_kind='bernoulli', h_kind='bernoulli')
    elif kind == 'gaussian':
        cfg.update(v_kind='gaussian', h_kind='bernoulli',
                   sigma=np.linspace(0.5, 1.5, V))
    elif kind == 'multinomial':
        cfg.update(v_kind='bernoulli', h_kind='multinomial', h_n_samples=20)
    cfg.update(kw)
    return cfg


def make_pair(cfg, seed=0):
    rng = np.random.RandomState(seed)
    V, H = cfg['n_visible'], cfg['n_hidden']
    dt = np.dtype(cfg['dtype'])
    init = dict(W=(0.1 * rng.randn(V, H)).astype(dt), vb=(0.1 * rng.randn(V)).astype(dt),
                hb=(0.1 * rng.randn(H)).astype(dt))
    eng, ora = _native.CudaRBM(cfg), OracleRBM(cfg)
    eng.set_params(init), ora.set_params(init)
    return eng, ora


def make_data(cfg, B, seed=1):
    rng = np.random.RandomState(seed)
    V = cfg['n_visible']
    if cfg['v_kind'] == 'gaussian':
        return rng.randn(B, V).astype(cfg['dtype'])
    return (rng.rand(B, V) < 0.3).astype(cfg['dtype'])


def oracle_acts(ora, X, k, seed, tick):
    Xp = ora.prepare_input(X, seed, tick)
    h0_means, v_states, v_means, _, h_means = ora.chain(Xp, k, seed, tick)
    # recompute h0 states exactly as chain() drew them
    from oracle import philox as P
    if ora.cfg.get('sample_h', True):
        h0s = ora._sample_h(ora.means_h_given_v(Xp), seed, P.SITE_H0, 0, tick, 0)
    else:
        h0s = h0_means
    return dict(X=Xp, h0_means=h0_means, h0_states=h0s, v_means=v_means, v_states=v_states, h_means=h_means)


SHAPES = [(37, 29, 19), (784, 16, 32), (130, 70, 65)]


@pytest.mark.parametrize('kind', ['bernoulli', 'gaussian', 'multinomial'])
@pytest.mark.parametrize('V,H,B', SHAPES)
@pytest.mark.parametrize('dtype', ['float32', 'float64'])
def test_cd1_activations_and_update_match_oracle(kind, V, H, B, dtype):
    """BASELINE configs[0]-style check: one CD-1 step, every intermediate and every variable."""
    cfg = make_cfg(kind, V, H, B, dtype=dtype, dropout=0.9)
    eng, ora = make_pair(cfg)
    X = make_data(cfg, B)
    seed, tick = 0x1234567, 3
    want = oracle_acts(ora, X, 1, seed, tick)
    eng.train_step(X, 0.05, 0.5, 1, seed, tick)
    # float64 gaussian units still draw float32 Box-Muller noise (libm sin/cos/log differ by ulps)
    tol = 2e-5 if dtype == 'float32' else (5e-6 if kind == 'gaussian' else 1e-9)
    for name in ACTS:
        got = eng.get_activation(name, B)
        if name.endswith('states') and kind != 'gaussian' or (name == 'h0_states'):
            bad = np.mean(got != want[name])
            assert bad <= 2e-3, (name, bad)
        else:
            np.testing.assert_allclose(got, want[name], atol=tol * max(1.0, np.abs(want[name]).max()), err_msg=name)
    ora.train_step(X, 0.05, 0.5, 1, seed, tick)
    g, w = eng.get_params(), ora.get_params()
    for k in ('W', 'vb', 'hb', 'dW', 'dvb', 'dhb', 'q_means'):
        np.testing.assert_allclose(g[k], w[k], atol=5e-6 if (dtype == 'float32' or kind == 'gaussian') else 1e-10, err_msg=k)
    eng.close()


@pytest.mark.parametrize('kind', ['bernoulli', 'gaussian', 'multinomial'])
def test_metrics_match_oracle(kind):
    cfg = make_cfg(kind, 130, 70, 65, sample_v=False)
    eng, ora = make_pair(cfg)
    X = make_data(cfg, 65)
    names = ('l2_loss', 'msre', 'pll', 'free_energy') if kind != 'gaussian' else ('l2_loss', 'msre', 'free_energy')
    got = eng.metrics(X, 2, 77, 5, names)
    want = ora.metrics(X, 2, 77, 5, names)
    for n in names:
        assert got[n] == pytest.approx(want[n], rel=2e-4, abs=2e-4), n
    # metrics returned by a training step see the pre-update weights
    got = eng.train_step(X, 0.05, 0.5, 2, 78, 6, metrics=names)
    want = ora.train_step(X, 0.05, 0.5, 2, 78, 6, metrics=names)
    for n in names:
        assert got[n] == pytest.approx(want[n], rel=2e-4, abs=2e-4), n
    eng.close()


def test_cdk_trajectory_stays_with_oracle():
    """10 CD-3 steps, probabilities for visibles (no sampling noise amplification on v)."""
    cfg = make_cfg('bernoulli', 96, 48, 64, sample_v=False, sparsity_cost=0.0)
    eng, ora = make_pair(cfg)
    rng = np.random.RandomState(5)
    for it in range(10):
        X = (rng.rand(64, 96) < 0.25).astype(np.float32)
        eng.train_step(X, 0.05, 0.6, 3, 4242, it)
        ora.train_step(X, 0.05, 0.6, 3, 4242, it)
    g, w = eng.get_params(['W', 'hb', 'vb']), ora.get_params(['W', 'hb', 'vb'])
    for k in g:
        np.testing.assert_allclose(g[k], w[k], atol=2e-4, err_msg=k)
    eng.close()


def test_transform_matches_oracle_and_ragged_batches():
    cfg = make_cfg('bernoulli', 50, 20, 16, sample_v=False)
    eng, ora = make_pair(cfg)
    for rows in (1, 7, 16, 33):      # 33 > max_batch: workspaces grow on demand
        X = make_data(cfg, rows, seed=rows)
        np.testing.assert_allclose(eng.transform(X, 2, 9, rows), ora.transform(X, 2, 9, rows), atol=2e-5)
    eng.close()


def test_resident_dataset_equals_fed_batches():
    cfg = make_cfg('bernoulli', 64, 32, 16, sample_v=False)
    a, _ = make_pair(cfg)
    b, _ = make_pair(cfg)
    X = make_data(cfg, 64)
    b.set_data(X)
    for it in range(4):
        a.train_step(X[it * 16:(it + 1) * 16], 0.05, 0.5, 1, 11, it)
        b.train_step_at(it * 16, 16, 0.05, 0.5, 1, 11, it)
    for k, v in a.get_params().items():
        np.testing.assert_array_equal(v, b.get_params()[k])
    a.close(), b.close()


@pytest.mark.parametrize('compute', ['fp32', 'bf16'])
@pytest.mark.parametrize('pinned', [False, True])
def test_train_epoch_equals_the_per_batch_loop(compute, pinned):
    """bm_rbm_train_epoch (double-buffered uploads, deferred metric read-back) is the same
    computation as calling bm_rbm_train_step per batch, ragged last batch included."""
    cfg = make_cfg('bernoulli', 96, 48, 16, compute=compute, sample_v=False, dropout=0.8)
    a, _ = make_pair(cfg)
    b, _ = make_pair(cfg)
    X = make_data(cfg, 16 * 5 + 7)
    Xh = _native.pinned_copy(X) if pinned else X
    want = []
    for i, lo in enumerate(range(0, len(X), 16)):
        m = a.train_step(X[lo:lo + 16], 0.05, 0.5, 2, 11, 3 + i, metrics=('msre', 'pll') if (4 + i + 1) % 2 == 0 else ())
        if m:
            want.append(m)
    got = b.train_epoch(Xh, 16, 0.05, 0.5, 2, 11, 3, metrics=('msre', 'pll'), every=2, iter0=4)
    assert len(got['msre']) == len(want) == 3
    np.testing.assert_allclose(got['msre'], [m['msre'] for m in want], rtol=1e-12)
    np.testing.assert_allclose(got['pll'], [m['pll'] for m in want], rtol=1e-12)
    for k, v in a.get_params().items():
        np.testing.assert_array_equal(v, b.get_params()[k])
    if pinned:
        _native.pinned_free(Xh)
    a.close(), b.close()


@pytest.mark.parametrize('compute,dtype', [('fp32', 'float32'), ('bf16', 'float32'), ('fp32', 'float64')])
@pytest.mark.parametrize('dropout', [None, 0.8])
def test_byte_valued_epoch_is_bit_identical_to_the_float_epoch(compute, dtype, dropout):
    """bm_rbm_train_epoch_u8 (one byte per visible unit, widened on the device) = bm_rbm_train_epoch on
    float(X): same parameters bit for bit, same metrics; V and the last batch deliberately ragged."""
    kw = dict(compute=compute, dtype=dtype, sample_v=False)
    if dropout:
        kw['dropout'] = dropout
    cfg = make_cfg('bernoulli', 100, 48, 16, **kw)
    a, _ = make_pair(cfg)
    b, _ = make_pair(cfg)
    X = make_data(cfg, 16 * 4 + 5)
    Xb = _native.as_bytes(X)
    assert Xb is not None and Xb.dtype == np.uint8
    P = _native.pinned_copy(Xb)
    want = a.train_epoch(X, 16, 0.05, 0.5, 2, 11, 3, metrics=('msre', 'pll', 'free_energy'), every=2, iter0=0)
    got = b.train_epoch(P, 16, 0.05, 0.5, 2, 11, 3, metrics=('msre', 'pll', 'free_energy'), every=2, iter0=0)
    for m in want:
        assert len(got[m]) == len(want[m]) == 2
        np.testing.assert_allclose(got[m], want[m], rtol=1e-12, err_msg=m)
    for k, v in a.get_params().items():
        np.testing.assert_array_equal(v, b.get_params()[k], err_msg=k)
    _native.pinned_free(P)
    a.close(), b.close()


def test_bfloat16_feed_epoch_is_bit_identical_to_the_float_epoch():
    """bm_rbm_train_epoch_bf16 (real-valued data packed as bfloat16 by bm_host_pack_bf16) = bm_rbm_train_epoch on the float32
    rows for the bf16 engine, which rounds its input to bfloat16 before the first GEMM: same parameters bit for bit; MSRE too
    (it is computed from the bf16 activations); refused by the fp32 engine and with dropout."""
    cfg = make_cfg('bernoulli', 100, 48, 16, compute='bf16', dtype='float32', sample_v=False)
    a, _ = make_pair(cfg)
    b, _ = make_pair(cfg)
    X = (np.random.RandomState(3).rand(16 * 4 + 5, 100) ** 2).astype(np.float32)          # grey levels in [0, 1)
    P = b.pin(X)
    assert isinstance(P, _native.Bf16Array) and P.shape == X.shape
    np.testing.assert_array_equal(P.widen(), np.asarray(_native.as_bf16(X)).view(_native.Bf16Array).widen())
    assert np.max(np.abs(P.widen() - X)) <= 2.0 ** -8
    want = a.train_epoch(X, 16, 0.05, 0.5, 2, 11, 3, metrics=('msre',), every=2, iter0=0)
    got = b.train_epoch(P, 16, 0.05, 0.5, 2, 11, 3, metrics=('msre',), every=2, iter0=0)
    np.testing.assert_allclose(got['msre'], want['msre'], rtol=1e-12)
    for k, v in a.get_params().items():
        np.testing.assert_array_equal(v, b.get_params()[k], err_msg=k)
    c, _ = make_pair(make_cfg('bernoulli', 100, 48, 16, compute='fp32', dtype='float32', sample_v=False))
    assert not isinstance(c.pin(X), _native.Bf16Array)               # (leaks one small pinned block; fine in a test)
    with pytest.raises(RuntimeError, match='bfloat16 feed'):
        c.train_epoch(P, 16, 0.05, 0.5, 2, 11, 3)
    b.unpin(P)
    a.close(), b.close(), c.close()


def test_fit_takes_the_byte_path_for_binary_data_and_matches_float_feeding(workdir, monkeypatch):
    """BaseRBM.fit pins binary data as bytes (engine.pin -> as_bytes); the trained weights equal those of a fit
    whose data went through the float32 path (the reference's feed_dict values, base_rbm.py:549-571)."""
    from boltzmann_machines.base import set_engine_factory
    from boltzmann_machines.rbm import BernoulliRBM
    old = set_engine_factory('rbm', None)
    try:
        rng = np.random.RandomState(5)
        X = (rng.rand(200, 64) < 0.3).astype(np.float32)
        seen = []
        real_empty = _native.pinned_empty
        monkeypatch.setattr(_native, 'pinned_empty', lambda shape, dtype=np.float32: seen.append(np.dtype(dtype)) or real_empty(shape, dtype))

        def fit(path):
            m = BernoulliRBM(n_visible=64, n_hidden=32, batch_size=32, max_epoch=2, random_seed=7, verbose=False,
                             sample_v_states=False, model_path=path,
                             metrics_config=dict(msre=True, train_metrics_every_iter=2))
            m.fit(X)
            return m.get_tf_params('weights')['W']
        W1 = fit('a/')
        assert seen and seen[-1] == np.uint8
        monkeypatch.setattr(_native, 'as_bytes', lambda Z, out=None: None)          # decline the byte path
        W2 = fit('b/')
        # the default (bf16) engine then takes real-valued data as bfloat16 -- exact for these values; BM_COMPUTE=fp32: float32
        assert seen[-1] in (np.uint16, np.float32)
        np.testing.assert_array_equal(W1, W2)
        assert np.abs(W1).max() > 0
    finally:
        set_engine_factory('rbm', old)


def test_init_weights_matches_tf_stream():
    from oracle import philox as P
    for dtype in ('float32', 'float64'):
        eng = _native.CudaRBM(make_cfg('bernoulli', 12, 8, 4, dtype=dtype))
        eng.init_normal_W(0.01, 1337)
        W = eng.get_params(['W'])['W']
        np.testing.assert_allclose(W, P.tf_random_normal((12, 8), 0.01, 1337, dtype), atol=1e-8)
        np.testing.assert_almost_equal(W[0][0], -0.0094548017 if dtype == 'float32' else -0.0077341544416)
        eng.close()


def test_bad_arguments_are_errors_not_crashes():
    eng = _native.CudaRBM(make_cfg('bernoulli', 12, 8, 4))
    with pytest.raises(ValueError):
        eng.train_step(np.zeros((4, 11), np.float32), 0.1, 0.5, 1, 1, 0)
    with pytest.raises(RuntimeError):
        eng.train_step(np.zeros((4, 12), np.float32), 0.1, 0.5, 0, 1, 0)      # k = 0
    with pytest.raises(RuntimeError):
        eng.train_step_at(0, 4, 0.1, 0.5, 1, 1, 0)                            # no resident data
    with pytest.raises(KeyError):
        eng.set_params({'nope': np.zeros(3)})
    eng.close()


@pytest.mark.parametrize('compute', ['fp32', 'bf16'])
def test_resident_dataset_taller_than_the_grid_limit(compute):
    """A resident dataset with more than 65535 rows (the bench holds 163840): the conversion kernels index rows with
    grid.y and must go in slabs -- a step on rows beyond the limit equals the same step fed from the host."""
    V, H, B, n = 24, 16, 64, 70000
    cfg = make_cfg('bernoulli', V, H, B, compute=compute, sample_v=False)
    rng = np.random.RandomState(3)
    X = (rng.rand(n, V) < 0.3).astype(np.float32)
    init = dict(W=(0.1 * rng.randn(V, H)).astype(np.float32))
    a, b = _native.CudaRBM(cfg), _native.CudaRBM(cfg)
    a.set_params(init), b.set_params(init)
    a.set_data(X)
    first = 69000
    a.train_step_at(first, B, 0.05, 0.5, 2, 11, 0)
    b.train_step(X[first:first + B], 0.05, 0.5, 2, 11, 0)
    ga, gb = a.get_params(), b.get_params()
    for k in ga:
        np.testing.assert_array_equal(ga[k], gb[k], err_msg=k)
    assert np.abs(ga['dW']).max() > 0
    a.close(); b.close()


@pytest.mark.parametrize('feed', ['bytes', 'bf16'])
def test_epochs_of_changing_batch_size_interleaved_with_single_steps(feed):
    """The bf16 engine's operand buffer has two halves (the copy stream converts batch i+1 while batch i runs) and grows with the
    largest batch seen: epochs with batch 16, then 40 (growth inside the epoch call), a single step, batch 16 again (halves now 40
    rows apart), a ragged epoch shorter than one batch -- bit-identical to an engine that is only ever fed batch by batch."""
    cfg = make_cfg('bernoulli', 100, 48, 64, compute='bf16', sample_v=False)
    a, _ = make_pair(cfg)
    b, _ = make_pair(cfg)
    rng = np.random.RandomState(3)
    if feed == 'bytes':
        X = (rng.rand(200, 100) < 0.3).astype(np.float32)
    else:
        X = rng.rand(200, 100).astype(np.float32)
    P = b.pin(X)
    assert (P.dtype == np.uint8) if feed == 'bytes' else isinstance(P, _native.Bf16Array)
    tick = 0
    for batch, rows in ((16, 200), (40, 200), (0, 33), (16, 70), (64, 9), (40, 81)):
        if batch == 0:                                        # a single step on rows of the float array
            a.train_step(X[:rows], 0.05, 0.5, 2, 11, tick)
            b.train_step(X[:rows], 0.05, 0.5, 2, 11, tick)
            tick += 1
            continue
        want = []
        for i, lo in enumerate(range(0, rows, batch)):
            want.append(a.train_step(X[lo:min(lo + batch, rows)], 0.05, 0.5, 2, 11, tick + i, metrics=('msre',))['msre'])
        got = b.train_epoch(P[:rows], batch, 0.05, 0.5, 2, 11, tick, metrics=('msre',), every=1)
        tick += len(want)
        np.testing.assert_allclose(got['msre'], want, rtol=1e-12)
        for k, v in a.get_params().items():
            np.testing.assert_array_equal(v, b.get_params()[k], err_msg='{0} after batch {1}'.format(k, batch))
    b.unpin(P)
    a.close(), b.close()
