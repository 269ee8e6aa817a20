"""The host mirror's utilities against golden vectors produced by the REFERENCE's own modules
(tests/golden/make_golden.py imports yell/boltzmann-machines' utils/utils.py and utils/rng.py unmodified and
records their outputs; the reference tree is not needed to run this test)."""
import json
import os

import numpy as np
import pytest

from boltzmann_machines.base.mixin import SeedMixin
from boltzmann_machines.utils import RNG
from boltzmann_machines.utils import utils as U

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_utils.json')))


def test_log_domain_statistics_match_the_reference():
    """log_sum_exp / log_mean_exp / log_std_exp / log_diff_exp post-process the AIS log-weights (dbm.py:843-870)."""
    rng = np.random.RandomState(20260923)           # the generator's input stream, replayed
    for rec in GOLD['log_stats']:
        n, loc, scale = rec['x_seed']
        x = (loc + scale * rng.randn(int(n))).astype(np.float64)
        if rec['x'] is not None:
            np.testing.assert_array_equal(x, np.asarray(rec['x']))
        assert float(U.log_sum_exp(x)) == pytest.approx(rec['log_sum_exp'], rel=1e-13, abs=1e-12)
        assert float(U.log_mean_exp(x)) == pytest.approx(rec['log_mean_exp'], rel=1e-13, abs=1e-12)
        if 'log_std_exp' in rec:
            assert float(U.log_std_exp(x)) == pytest.approx(rec['log_std_exp'], rel=1e-9)
            np.testing.assert_allclose(U.log_diff_exp(np.sort(x))[:8], rec['log_diff_exp_sorted_head'], rtol=1e-9)


def test_iteration_helpers_match_the_reference():
    for rec in GOLD['batch_iter']:
        X = np.arange(rec['n'] * 2).reshape(rec['n'], 2)
        assert [b.tolist() for b in U.batch_iter(X, batch_size=rec['batch_size'])] == rec['batches']
        assert [X[lo:hi].tolist() for lo, hi in U.batch_bounds(rec['n'], rec['batch_size'])] == rec['batches']
    for rec in GOLD['epoch_iter']:
        assert [int(e) for e in U.epoch_iter(rec['start'], rec['max'])] == rec['epochs']
    for rec in GOLD['make_list_from']:
        v = tuple(rec['in']) if rec.get('tuple') else rec['in']
        assert list(U.make_list_from(v)) == rec['out']
    for rec in GOLD['one_hot']:
        assert np.asarray(U.one_hot(rec['y'])).tolist() == rec['one_hot']
        assert np.asarray(U.unhot(U.one_hot(rec['y']))).tolist() == rec['unhot']


def test_host_rng_matches_the_reference():
    """random_state.json interchangeability and the per-call seeds of run_in_tf_session (base/mixin.py:34-35)."""
    for rec in GOLD['rng']:
        g = RNG(seed=rec['seed'])
        assert [float(v) for v in g.rand(3)] == rec['rand3']
        assert [int(g.randint(2 ** 31 - 1)) for _ in range(3)] == rec['make_random_seed']
        state = g.get_state()
        assert float(g.rand()) == rec['next_after_state']
        assert float(RNG(seed=None).set_state(json.loads(json.dumps(state))).rand()) == rec['replayed_after_state']
        assert float(g.reseed().rand()) == rec['reseeded_first']

        class M(SeedMixin):
            pass
        m = M(random_seed=rec['seed'])
        m._rng.rand(3)
        assert [m.make_random_seed() for _ in range(3)] == rec['make_random_seed']


def test_layer_plugins_match_the_reference_layers():
    """layers.py plug-in surface (north_star): `activation(x, b)` of the three unit types in both dtypes and the shape /
    dtype contract of `init`, against the reference's own layers.py evaluated on the shim
    (tests/golden/reference_layers.json, written by make_reference_golden.py)."""
    import json
    import os
    import numpy as np
    from boltzmann_machines import layers as L
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_layers.json')
    cases = json.load(open(path))['cases']
    assert len(cases) == 6
    for c in cases:
        n = len(c['b'])
        kw = dict(n_units=n, dtype=c['dtype'])
        if c['cls'] == 'MultinomialLayer':
            kw['n_samples'] = 7
        if c['cls'] == 'GaussianLayer':
            kw['sigma'] = np.linspace(0.5, 1.5, n)
        layer = getattr(L, c['cls'])(**kw)
        x, b = np.asarray(c['x'], dtype=c['dtype']), np.asarray(c['b'], dtype=c['dtype'])
        got = np.asarray(layer.activation(x, b))
        np.testing.assert_allclose(got, np.asarray(c['activation']), rtol=2e-6 if c['dtype'] == 'float32' else 1e-12,
                                   atol=1e-7 if c['dtype'] == 'float32' else 1e-14, err_msg=c['cls'] + ' ' + c['dtype'])
        init = layer.init(batch_size=4, random_seed=5)
        assert list(init.shape) == c['init_shape'] and str(init.dtype) == c['init_dtype'], c['cls']
        s = layer.sample(got)
        assert s.shape == got.shape and str(s.dtype) == c['dtype']
