"""BASELINE.json's parity gates for the bf16 tensor-core mode, evaluated on the CPU at the benchmark's width: the oracle
that rounds operands to bf16 where the engine does (oracle/rbm.py, compute='bf16') against the storage-precision oracle,
trained side by side on the same synthetic 'MNIST' with the same Philox streams (BernoulliRBM 784-1024, CD-5).

north_star: pseudo-log-likelihood within +-0.5 nats; SURVEY 8d: MSRE within 2 %."""
import numpy as np

from oracle.rbm import OracleRBM


def teacher_data(n, V=784, seed=1337):
    rng = np.random.RandomState(seed)
    Wt = (0.5 * rng.randn(V, 64)).astype(np.float32)
    bt = np.float32(np.log(0.13 / 0.87))
    v = (rng.rand(n, V) < 0.5).astype(np.float32)
    for _ in range(20):
        h = (rng.rand(n, 64) < 1.0 / (1.0 + np.exp(-(v @ Wt)))).astype(np.float32)
        v = (rng.rand(n, V) < 1.0 / (1.0 + np.exp(-(h @ Wt.T + bt)))).astype(np.float32)
    return v


def test_bf16_mode_keeps_pll_within_half_a_nat_and_msre_within_two_percent():
    V, H, B, k, steps = 784, 1024, 256, 5, 24
    X = teacher_data(B * 4 + 512)
    X_train, X_val = X[:B * 4], X[B * 4:]
    rng = np.random.RandomState(0)
    W0 = (0.01 * rng.randn(V, H)).astype(np.float32)
    p = np.clip(X_train.mean(axis=0), 1e-3, 1 - 1e-3)
    vb0 = np.log(p / (1 - p)).astype(np.float32)
    engines = []
    for compute in ('fp32', 'bf16'):
        cfg = dict(n_visible=V, n_hidden=H, v_kind='bernoulli', h_kind='bernoulli', dtype='float32', compute=compute,
                   l2=1e-5, sample_v=False, sample_h=True, max_batch=B)
        e = OracleRBM(cfg)
        e.set_params({'W': W0, 'vb': vb0})
        engines.append(e)
    for it in range(steps):
        lo = (it % 4) * B
        for e in engines:
            e.train_step(X_train[lo:lo + B], 0.05, 0.5 if it < 12 else 0.8, k, 20260923, it)
    ref, emu = engines
    Wa, Wb = ref.get_params(['W'])['W'], emu.get_params(['W'])['W']
    assert np.linalg.norm(Wa - Wb) / np.linalg.norm(Wa) < 0.05
    ma = ref.metrics(X_val, k, 7, 1000, ('msre', 'pll', 'free_energy'))
    mb = emu.metrics(X_val, k, 7, 1000, ('msre', 'pll', 'free_energy'))
    assert abs(ma['pll'] - mb['pll']) < 0.5, (ma['pll'], mb['pll'])
    assert abs(ma['msre'] - mb['msre']) < 0.02 * ma['msre'], (ma['msre'], mb['msre'])
    assert abs(ma['free_energy'] - mb['free_energy']) < 0.5, (ma['free_energy'], mb['free_energy'])
