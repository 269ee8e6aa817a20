"""The host mirror + oracle (CPU) and the host mirror + CUDA engine (GPU, fp32 compute) against golden vectors
produced by the REFERENCE'S OWN RBM code (tests/golden/make_reference_rbm_golden.py: yell/boltzmann-machines imported
unmodified, TensorFlow replaced by oracle/tf1shim.py, random ops answered from the shared Philox layout).

Each case replays a whole public-API scenario -- fit() with schedules / validation metrics / free-energy gap,
transform(), get_tf_params() -- so the comparison covers gradients, sparsity, momentum, the metrics' exact forms,
per-call seeding and tick order, and persistence between calls.  The reference tree is not needed to run this test.
(The file name sorts last on purpose: the GPU variant could not be run before the end of round 1.)"""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_rbm_cases.json')))
CASES = {c['name']: c for c in GOLD['cases']}


@pytest.fixture(params=['oracle', pytest.param('cuda-fp32', marks=pytest.mark.gpu)])
def engine_kind(request, monkeypatch):
    from boltzmann_machines.base import set_engine_factory
    if request.param == 'oracle':
        from oracle.rbm import rbm_factory
        old = set_engine_factory('rbm', rbm_factory)
    else:
        monkeypatch.setenv('BM_COMPUTE', 'fp32')
        old = set_engine_factory('rbm', None)
    yield request.param
    set_engine_factory('rbm', old)


def build(case, workdir):
    from boltzmann_machines import rbm as R
    kw = dict(case['kw'])
    dt = kw.get('dtype', 'float32')
    for k in ('W_init',):
        if isinstance(kw.get(k), list):
            kw[k] = np.asarray(kw[k], dtype=dt)
    kw['model_path'] = os.path.join(str(workdir), case['name']) + '/'
    model = getattr(R, case['cls'])(**kw)
    log = {'train': [], 'val': [], 'feg': []}
    for meth, key in (('_train_epoch', 'train'), ('_run_val_metrics', 'val'), ('_run_feg', 'feg')):
        orig = getattr(model, meth)

        def wrapped(*a, _orig=orig, _key=key, **k):
            r = _orig(*a, **k)
            log[_key].append(r)
            return r
        setattr(model, meth, wrapped)
    return model, log, dt


def close(got, want, tol, what):
    if want is None:
        assert got is None, what
        return
    np.testing.assert_allclose(np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64),
                               rtol=tol, atol=tol, err_msg=what)


@pytest.mark.parametrize('name', sorted(CASES))
def test_public_api_scenario_matches_the_reference(name, engine_kind, workdir):
    case = CASES[name]
    model, log, dt = build(case, workdir)
    # float32: the same formulas in float32 with different summation orders; a Bernoulli draw is u < p on the SAME u
    tol = 1e-9 if dt == 'float64' else (2e-5 if engine_kind == 'oracle' else 2e-4)
    if case['X'] is None:
        model.init()
    else:
        X = np.asarray(case['X'], dtype=dt)
        X_val = None if case['X_val'] is None else np.asarray(case['X_val'], dtype=dt)
        model.fit(X, X_val)
        H = model.transform(X[:case['transform_rows']])
        close(H, case['transform'], tol * 5, 'transform')
    assert (int(model.epoch_), int(model.iter_)) == (case['epoch_'], case['iter_'])
    for scope in ('weights', 'grads_accumulators'):
        got = model.get_tf_params(scope=scope)
        assert sorted(got) == sorted(case[scope]), scope
        for k, want in case[scope].items():
            close(got[k], want, tol, '{0}/{1}'.format(scope, k))
    if case['X'] is None:
        return
    # metrics: per-epoch means of the reporting iterations, validation metrics, free-energy gap
    mtol = dict(msre=10 * tol, l2_loss=10 * tol, pll=2e-3 if dt == 'float32' else 1e-7)
    assert len(log['train']) == len(case['log']['train'])
    for got, want in zip(log['train'], case['log']['train']):
        for m, v in want.items():
            close(got.get(m), v, mtol[m], 'train ' + m)
    assert len(log['val']) == len(case['log']['val'])
    for got, want in zip(log['val'], case['log']['val']):
        for m, v in want.items():
            close(got.get(m), v, mtol[m], 'val ' + m)
    # a free energy is a sum over ~V+H terms of magnitude ~10: float32 rounding of the batch means dominates
    close(log['feg'], case['log']['feg'], 5e-4 if dt == 'float32' else 1e-8, 'feg')
