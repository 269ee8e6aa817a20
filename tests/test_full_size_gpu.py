"""Parity at BASELINE.json's FULL sizes.

A k-step Gibbs chain is chaotic in its samples (one Bernoulli draw that flips at rounding level changes
everything downstream), so at full size the CUDA path is checked through properties that do not depend on
the size or on the sample path, each evaluated by the oracle's own functions on the engine's OWN activations:

  (a) every conditional that can be recomputed from exported activations -- h0 | X, v_1 | h0 (k = 1),
      h_k | v_k -- equals the oracle's (operands rounded to bf16 at the same points) up to one bf16 ulp;
  (b) sampled states are exactly the shared Philox stream's (u < p) (only draws within a bf16 ulp of p may differ);
  (c) the parameter update equals the oracle's gradients + sparsity + momentum formulas applied to those
      activations (statistics kernels, split-K dW, fused reduce + update);
  (d) aggregate statistics (MSRE, mean activation) agree with an independent oracle run of the same step.

cfg2: BernoulliRBM 784-1024, batch 4096, CD-5.   cfg3: GaussianRBM 3072-5000, batch 2048, CD-1.
cfg5 (per-GPU shard): BernoulliRBM 784-4096, batch 4096, 25 Gibbs steps.   cfg4: DBM 784-512-1024, batch = particles = 1024,
25 mean-field updates -- small enough for the oracle to run whole, compared directly.
"""
import numpy as np
import pytest

from boltzmann_machines import _native
from oracle import philox as P
from oracle.rbm import OracleRBM
from oracle.dbm import OracleDBM

pytestmark = pytest.mark.gpu


def bern_data(rows, V, seed):
    """Binary rows with a 13 % on-rate and some structure (a few prototypes + flips)."""
    rng = np.random.RandomState(seed)
    proto = rng.rand(16, V) < 0.13
    X = proto[rng.randint(0, 16, size=rows)] ^ (rng.rand(rows, V) < 0.03)
    return X.astype(np.float32)


def one_ulp(got, want, name, frac=0.05):
    np.testing.assert_allclose(got, want, rtol=2.0 ** -7, atol=2e-6, err_msg=name)
    assert np.mean(got != want) <= frac, (name, float(np.mean(got != want)))


CASES = {
    'cfg2-cd1': dict(kind='bernoulli', V=784, H=1024, B=4096, k=1, w_std=0.01),
    'cfg2-cd5': dict(kind='bernoulli', V=784, H=1024, B=4096, k=5, w_std=0.01),
    'cfg3-gauss-cd1': dict(kind='gaussian', V=3072, H=5000, B=2048, k=1, w_std=0.0008),
    'cfg5-shard-k25': dict(kind='bernoulli', V=784, H=4096, B=4096, k=25, w_std=0.01),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_full_size_step_properties(name):
    c = CASES[name]
    V, H, B, k = c['V'], c['H'], c['B'], c['k']
    cfg = dict(n_visible=V, n_hidden=H, dtype='float32', compute='bf16', l2=1e-5, max_batch=B,
               sample_v=False, sample_h=True, sparsity_cost=0.01, sparsity_target=0.1, sparsity_damping=0.9,
               v_kind=c['kind'], h_kind='bernoulli')
    rng = np.random.RandomState(11)
    if c['kind'] == 'gaussian':
        cfg['sigma'] = np.ones(V)
        Z = rng.randn(B, 64).astype(np.float32)
        X = (Z @ rng.randn(64, V).astype(np.float32) + 0.3 * rng.randn(B, V).astype(np.float32))
        X = ((X - X.mean(0)) / X.std(0)).astype(np.float32)
        lr = 5e-4
    else:
        X = bern_data(B, V, 3)
        lr = 0.05
    init = dict(W=(c['w_std'] * rng.randn(V, H)).astype(np.float32), vb=(0.1 * rng.randn(V)).astype(np.float32),
                hb=(0.1 * rng.randn(H)).astype(np.float32), dW=(1e-3 * rng.randn(V, H)).astype(np.float32))
    eng, ora, ref = _native.CudaRBM(cfg), OracleRBM(cfg), OracleRBM(cfg)
    for e in (eng, ora, ref):
        e.set_params(init)
    seed, tick, mom = 0x5EED, 4, 0.5

    got_m = eng.train_step(X, lr, mom, k, seed, tick, metrics=('msre',))
    A = {n: eng.get_activation(n, B) for n in ('X', 'h0_means', 'h0_states', 'v_means', 'v_states', 'h_means')}

    # (a) conditionals recomputed by the oracle from the engine's own activations (pre-update weights)
    Xp = ora.prepare_input(X, seed, tick)
    np.testing.assert_array_equal(A['X'], Xp)
    one_ulp(A['h0_means'], ora._r(ora.means_h_given_v(Xp)), 'h0_means')
    if k == 1:
        one_ulp(A['v_means'], ora._r(ora.means_v_given_h(A['h0_states'])), 'v_means')
    np.testing.assert_array_equal(A['v_states'], A['v_means'])           # sample_v = False
    one_ulp(A['h_means'], ora._r(ora.means_h_given_v(A['v_states'])), 'h_means')

    # (b) the Bernoulli draws are the shared Philox stream's
    u = P.uniform_at(B, H, seed, P.SITE_H0, 0, tick)
    p, s = A['h0_means'], A['h0_states']
    assert set(np.unique(s)) <= {0.0, 1.0}
    disagree = s != (u < p)
    assert disagree.mean() < 0.01
    assert np.all(np.abs(u - p)[disagree] <= 2.0 ** -8 * np.maximum(p[disagree], 1e-3))

    # (c) update = the oracle's formulas on these activations
    ora.apply_update(A['X'], A['h0_means'], A['v_states'], A['h_means'], lr, mom)
    g, w = eng.get_params(), ora.get_params()
    for n in ('dW', 'W', 'dvb', 'vb', 'dhb', 'hb', 'q_means'):
        scale = max(1e-6, float(np.abs(w[n]).max()))
        np.testing.assert_allclose(g[n], w[n], atol=2e-5 * scale + 1e-7, rtol=1e-5, err_msg=n)

    # (d) an independent oracle run of the same step agrees in aggregate
    want_m = ref.train_step(X, lr, mom, k, seed, tick, metrics=('msre',))
    assert got_m['msre'] == pytest.approx(want_m['msre'], rel=0.02)
    rw = ref.get_params()
    assert np.abs(g['hb'] - rw['hb']).max() < 5e-3 * max(lr / 0.05, 1e-2) + 1e-5
    if k <= 5:                                   # longer chains decorrelate sample by sample; aggregates above still hold
        rel = np.linalg.norm(g['W'] - rw['W']) / np.linalg.norm(rw['W'] - init['W'])
        assert rel < 0.5, rel                    # the two updates point the same way
    eng.close()


def test_full_size_resident_epoch_is_deterministic_and_matches_fed_batches():
    """cfg2 through the three entry points the bench uses: dataset-resident steps, float32 epochs and
    byte-valued epochs give bit-identical parameters."""
    V, H, B, k = 784, 1024, 4096, 5
    cfg = dict(n_visible=V, n_hidden=H, dtype='float32', compute='bf16', l2=1e-5, max_batch=B,
               sample_v=False, sample_h=True, v_kind='bernoulli', h_kind='bernoulli')
    X = bern_data(3 * B, V, 5)
    out = []
    for mode in ('resident', 'epoch_f32', 'epoch_u8'):
        eng = _native.CudaRBM(cfg)
        eng.init_normal_W(0.01, 1337)
        if mode == 'resident':
            eng.set_data(X)
            for i in range(3):
                eng.train_step_at(i * B, B, 0.05, 0.5, k, 77, i)
        else:
            Xh = eng.pin(X) if mode == 'epoch_u8' else _native.pinned_copy(X)
            assert Xh.dtype == (np.uint8 if mode == 'epoch_u8' else np.float32)
            m = eng.train_epoch(Xh, B, 0.05, 0.5, k, 77, 0, metrics=('msre',), every=1)
            assert len(m['msre']) == 3 and all(0.0 < v < 0.5 for v in m["msre"])
            _native.pinned_free(Xh)
        out.append(eng.get_params())
        eng.close()
    for n in out[0]:
        np.testing.assert_array_equal(out[0][n], out[1][n], err_msg=n)
        np.testing.assert_array_equal(out[0][n], out[2][n], err_msg=n)


def test_full_size_dbm_step_matches_the_oracle():
    """cfg4: DBM 784-512-1024, batch = particles = 1024, 25 mean-field updates, one PCD step -- the oracle
    runs the whole step in a few seconds, so this is a direct comparison (fp32 CUDA-core path)."""
    V, Hs, B = 784, [512, 1024], 1024
    cfg = dict(n_visible=V, n_hiddens=Hs, v_kind='bernoulli', h_kinds=['bernoulli'] * 2, h_n_samples=[100.] * 2,
               dtype='float32', compute='fp32', n_particles=B, batch_size=B, max_mf_updates=25, mf_tol=1e-7, l2=1e-7, max_norm=6.0,
               sample_v=True, sample_h=[True, True], sparsity_target=[0.2, 0.1], sparsity_cost=[1e-4, 5e-5],
               sparsity_damping=0.9)
    rng = np.random.RandomState(2)
    eng, ora = _native.CudaDBM(cfg), OracleDBM(cfg)
    d = {'vb': (0.1 * rng.randn(V)).astype(np.float32),
         'W': (0.02 * rng.randn(V, Hs[0])).astype(np.float32), 'hb': (0.1 * rng.randn(Hs[0])).astype(np.float32),
         'W_1': (0.02 * rng.randn(Hs[0], Hs[1])).astype(np.float32), 'hb_1': (0.1 * rng.randn(Hs[1])).astype(np.float32)}
    for e in (eng, ora):
        e.set_params(d)
        e.init_particles(4242)
    X = bern_data(B, V, 9)
    got = eng.train_step(X, 2e-3, 0.5, 1, 99, 0, metrics=('msre', 'n_mf_updates'))
    want = ora.train_step(X, 2e-3, 0.5, 1, 99, 0, metrics=('msre', 'n_mf_updates'))
    assert got['n_mf_updates'] == want['n_mf_updates']
    assert got['msre'] == pytest.approx(want['msre'], rel=1e-3)
    g, w = eng.get_params(), ora.get_params()
    for n in ('W', 'W_1', 'vb', 'hb', 'hb_1'):
        np.testing.assert_allclose(g[n], w[n], atol=1e-5, err_msg=n)       # the update itself is ~1e-3 (a flipped particle bit: lr/M = 2e-6)
        assert np.abs(w[n] - d[n]).max() > 1e-6, n
    for n in ('v', 'h', 'h_1'):
        assert np.mean(g[n] != w[n]) < 2e-3, n                             # particles: same Philox stream
    eng.close()


def test_north_star_gates_at_cfg2_after_three_epochs():
    """BASELINE.json's parity gates evaluated at configs[1]'s own size: BernoulliRBM 784-1024, batch 4096, CD-5, three epochs
    of the synthetic set (45 steps, the example's momentum schedule start and learning rate).  Validation PLL of the bf16
    tensor-core engine within +-0.5 nats and MSRE within 2 % of (a) the fp32 CUDA-core engine (bit-comparable with the oracle,
    tests/test_rbm_gpu.py) after all three epochs and (b) the CPU oracle itself after the first epoch."""
    import bench
    V, H, B, k = bench.V, bench.H, bench.B, bench.K_GIBBS
    n_batches = 15
    X = bench.synth_mnist(B * n_batches + B, seed=4321)
    Xtr, Xval = X[:B * n_batches], X[B * n_batches:]
    p = np.clip(Xtr.mean(axis=0), 1e-7, 1 - 1e-7)
    init = {'W': (0.01 * np.random.RandomState(5).randn(V, H)).astype(np.float32), 'vb': np.log(p / (1 - p)).astype(np.float32)}
    engines = {c: _native.CudaRBM(bench.model_cfg(c)) for c in ('bf16', 'fp32')}
    ora = OracleRBM(bench.model_cfg('fp32'))
    for e in list(engines.values()) + [ora]:
        e.set_params(init)
    seed, tick = 777, 0
    after = {}
    for epoch in range(3):
        for name, e in engines.items():
            e.train_epoch(Xtr, B, bench.LR, bench.MOMENTUM, k, seed, tick)
        if epoch == 0:
            for i in range(n_batches):
                ora.train_step(Xtr[i * B:(i + 1) * B], bench.LR, bench.MOMENTUM, k, seed, tick + i)
            after['oracle'] = ora.metrics(Xval, 1, seed, 10 ** 6, ('msre', 'pll'))
            after['bf16@1'] = engines['bf16'].metrics(Xval, 1, seed, 10 ** 6, ('msre', 'pll'))
        tick += n_batches
    m = {name: e.metrics(Xval, 1, seed, 10 ** 6, ('msre', 'pll')) for name, e in engines.items()}
    # training moved the model (the untrained model's PLL is about -V log 2 = -543)
    assert m['fp32']['pll'] > -400.0, m
    assert abs(m['bf16']['pll'] - m['fp32']['pll']) < 0.5, m
    assert abs(m['bf16']['msre'] - m['fp32']['msre']) < 0.02 * m['fp32']['msre'], m
    assert abs(after['bf16@1']['pll'] - after['oracle']['pll']) < 0.5, after
    assert abs(after['bf16@1']['msre'] - after['oracle']['msre']) < 0.02 * after['oracle']['msre'], after
    for e in engines.values():
        e.close()
