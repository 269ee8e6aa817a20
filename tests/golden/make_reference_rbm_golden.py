"""Golden vectors from the REFERENCE'S OWN RBM CODE: tests/golden/reference_rbm_cases.json.

Run in the build container only (reads /root/reference):

    python tests/golden/make_reference_rbm_golden.py

yell/boltzmann-machines is imported unmodified from /root/reference with `oracle/tf1shim.py` registered as
`tensorflow` (TensorFlow 1.3 itself cannot be installed here).  For every case below the reference's own
`BernoulliRBM/GaussianRBM/MultinomialRBM(...).fit(X, X_val)`, `.transform(X)`, `.get_tf_params(...)` and `.init()` run:
its graph construction (base_rbm.py:244-525), its fit loop (:533-655), its persistence round trip between public calls
(tf_model.py:10-40,117-162).  Every random op of the graph asks the provider below, which answers from the Philox
layout the oracle and the CUDA engine share (oracle/philox.py): draw site from the op's name scope, Gibbs index from
the `gibbs_step[_i]` scope (or the while_loop iteration), tick = index of the Session.run call inside the public call,
seed = the graph-level seed the reference sets from its own RNG (`tf.set_random_seed(model.make_random_seed())`).

What the golden file pins: the oracle's (and, on the GPU, the engine's) gradients, sparsity term, momentum updates,
metrics (MSRE, PLL with its batch-mean quirk, L2, free-energy gap), schedules, dropout, the Gaussian / Multinomial unit
types and the fixed / variable-length Gibbs chain AGAINST THE REFERENCE'S SOURCE executing those formulas.  What it
does not pin: TensorFlow's own random stream (substituted, see above) and TensorFlow's kernels (numpy here).
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import philox as P                      # noqa: E402
from oracle import tf1shim                          # noqa: E402
from oracle.rbm import multinomial_counts           # noqa: E402

FALLBACK_SEED = 0            # seed of public calls that do not reseed the graph (init, get_tf_params)


def gibbs_index(req):
    """(is_chain_step, t): t-th Gibbs step (1-based) from the `gibbs_step[_i]` scope or the while_loop iteration."""
    for part in req.scope.split('/'):
        if part == 'gibbs_step' or part.startswith('gibbs_step_'):
            if req.loop_iter is not None:
                return True, req.loop_iter + 1
            return True, (int(part.split('_')[-1]) if part != 'gibbs_step' else 0) + 1
    return False, 0


def provider(req):
    seed = req.graph_seed if req.graph_seed is not None else FALLBACK_SEED
    tick = req.run_index
    shape = tuple(int(s) for s in req.shape)
    if req.kind == 'normal':                        # W initialiser: tf.random_normal(..., seed=random_seed)
        assert req.scope.startswith('weights') and req.op_seed is not None, req.scope
        return P.tf_random_normal(shape, req.stddev, int(req.op_seed), np.dtype(req.dtype).name)
    if req.kind == 'dropout_uniform':
        return P.uniform_at(shape[0], shape[1], seed, P.SITE_DROPOUT, 0, tick)
    if req.kind == 'uniform_int':
        assert 'pseudo_loglik' in req.scope
        return P.site_words(shape[0], 1, seed, P.SITE_PLL, 0, tick)[:, 0]
    in_step, t = gibbs_index(req)
    if 'free_energy' in req.scope:                  # MultinomialRBM._free_energy: h_hat ~ Multinomial(M, uniform)
        assert req.kind == 'multinomial'
        leaf = [p for p in req.scope.split('/') if p.startswith('free_energy')][0]
        fe_idx = 2 if not req.scope.startswith('pseudo_loglik') else (0 if leaf == 'free_energy' else 1)
        K = shape[-1]
        probs = np.full((1, K), 1.0 / K, dtype=np.float32)
        return multinomial_counts(probs, int(req.total_count), seed, P.SITE_MULTINOMIAL_FE, fe_idx, tick)[0].astype(np.float32)
    if 'sample_h_given_v' in req.scope:
        site = P.SITE_H if in_step else P.SITE_H0
    elif 'sample_v_given_h' in req.scope:
        site = P.SITE_V
    else:
        raise AssertionError('unmapped random op: {0} {1}'.format(req.kind, req.name))
    rows, n = shape
    if req.kind == 'bernoulli':
        p = np.asarray(req.args[0])
        u = P.uniform_at(rows, n, seed, site, t, tick)
        return ((u < p) if p.dtype == np.float32 else (u.astype(np.float64) < p)).astype(np.int32)
    if req.kind == 'multinomial':                   # tf.multinomial is shift-invariant per row
        p = np.asarray(req.args[0], dtype=np.float64)
        probs = (p / p.sum(axis=1, keepdims=True)).astype(np.float32)
        return multinomial_counts(probs, int(req.total_count), seed, site, t, tick).astype(np.float32)
    if req.kind == 'normal_loc_scale':
        loc, scale = np.asarray(req.args[0]), np.asarray(req.args[1])
        return (loc + scale * P.normal_at(rows, n, seed, site, t, tick).astype(loc.dtype)).astype(loc.dtype)
    raise AssertionError(req.kind)


def cases():
    rng = np.random.RandomState(7)
    Xb = (rng.rand(40, 20) < 0.3).astype(np.float32)
    Xb_val = (rng.rand(16, 20) < 0.3).astype(np.float32)
    Xg = rng.randn(36, 16).astype(np.float32)
    Xg_val = rng.randn(12, 16).astype(np.float32)
    mc = dict(msre=True, pll=True, l2_loss=True, feg=True, train_metrics_every_iter=2, val_metrics_every_epoch=1,
              feg_every_epoch=2, n_batches_for_feg=3)
    W20 = (0.1 * rng.randn(20, 12)).astype(np.float32)
    out = []
    out.append(dict(name='bernoulli_cd1_schedules', cls='BernoulliRBM', X=Xb, X_val=Xb_val, transform_rows=10,
                    kw=dict(n_visible=20, n_hidden=12, W_init=W20, vb_init=-0.5, hb_init=0.1, n_gibbs_steps=1,
                            learning_rate=[0.05, 0.02], momentum=[0.5, 0.9], max_epoch=3, batch_size=8, l2=1e-3,
                            sparsity_target=0.2, sparsity_cost=0.05, sparsity_damping=0.8, random_seed=1337,
                            metrics_config=dict(mc), verbose=False, save_after_each_epoch=True)))
    out.append(dict(name='bernoulli_variable_k_dropout_sample_v', cls='BernoulliRBM', X=Xb, X_val=Xb_val, transform_rows=8,
                    kw=dict(n_visible=20, n_hidden=12, W_init=W20, n_gibbs_steps=[1, 2, 3], learning_rate=0.03,
                            momentum=0.7, max_epoch=3, batch_size=16, l2=1e-4, sample_v_states=True, dropout=0.8,
                            dbm_first=True, random_seed=42, metrics_config=dict(mc), verbose=False,
                            save_after_each_epoch=False)))
    out.append(dict(name='bernoulli_fixed_k3_means_only', cls='BernoulliRBM', X=Xb, X_val=None, transform_rows=8,
                    kw=dict(n_visible=20, n_hidden=12, W_init=W20, n_gibbs_steps=3, learning_rate=0.05, momentum=0.5,
                            max_epoch=2, batch_size=10, l2=0., sample_h_states=False, dbm_last=True, random_seed=5,
                            metrics_config=dict(msre=True, train_metrics_every_iter=1), verbose=False,
                            save_after_each_epoch=False)))
    out.append(dict(name='bernoulli_float64', cls='BernoulliRBM', X=Xb.astype(np.float64), X_val=Xb_val.astype(np.float64),
                    transform_rows=6,
                    kw=dict(n_visible=20, n_hidden=12, W_init=W20.astype(np.float64), n_gibbs_steps=2, learning_rate=0.05,
                            momentum=0.6, max_epoch=2, batch_size=8, l2=1e-3, sparsity_cost=0.01, random_seed=11,
                            dtype='float64', metrics_config=dict(mc), verbose=False, save_after_each_epoch=False)))
    out.append(dict(name='gaussian_sigma_vector', cls='GaussianRBM', X=Xg, X_val=Xg_val, transform_rows=9,
                    kw=dict(n_visible=16, n_hidden=10, W_init=(0.05 * rng.randn(16, 10)).astype(np.float32),
                            sigma=np.linspace(0.6, 1.4, 16).tolist(), n_gibbs_steps=1, learning_rate=2e-3, momentum=0.8,
                            max_epoch=3, batch_size=12, l2=1e-3, sample_v_states=True, random_seed=3,
                            metrics_config=dict(mc), verbose=False, save_after_each_epoch=False)))
    out.append(dict(name='multinomial_hidden', cls='MultinomialRBM', X=Xb, X_val=Xb_val, transform_rows=7,
                    kw=dict(n_visible=20, n_hidden=6, n_samples=10, W_init=(0.1 * rng.randn(20, 6)).astype(np.float32),
                            n_gibbs_steps=1, learning_rate=0.01, momentum=0.5, max_epoch=2, batch_size=8, l2=1e-3,
                            random_seed=9, metrics_config=dict(mc), verbose=False, save_after_each_epoch=False)))
    out.append(dict(name='init_from_seed', cls='BernoulliRBM', X=None, X_val=None, transform_rows=0,
                    kw=dict(n_visible=20, n_hidden=12, W_init=0.01, vb_init=0.25, random_seed=1337, verbose=False)))
    return out


def tolist(a):
    return None if a is None else np.asarray(a).tolist()


def run_case(ref_rbm, case, workdir):
    cls = getattr(ref_rbm, case['cls'])
    kw = dict(case['kw'])
    kw['model_path'] = os.path.join(workdir, case['name']) + '/'
    model = cls(**kw)
    log = {'train': [], 'val': [], 'feg': []}
    for meth, key in (('_train_epoch', 'train'), ('_run_val_metrics', 'val'), ('_run_feg', 'feg')):
        orig = getattr(model, meth)

        def wrapped(*a, _orig=orig, _key=key, **k):
            r = _orig(*a, **k)
            log[_key].append(r if not isinstance(r, dict) else {m: (None if v is None else float(v)) for m, v in r.items()})
            return r
        setattr(model, meth, wrapped)
    rec = {'name': case['name'], 'cls': case['cls'], 'kw': {k: (tolist(v) if isinstance(v, np.ndarray) else v) for k, v in case['kw'].items()},
           'X': tolist(case['X']), 'X_val': tolist(case['X_val']), 'transform_rows': case['transform_rows']}
    if case['X'] is None:
        model.init()
    else:
        model.fit(case['X'], case['X_val'])
        rec['transform'] = tolist(model.transform(case['X'][:case['transform_rows']]))
    rec['weights'] = {k: tolist(v) for k, v in model.get_tf_params(scope='weights').items()}
    rec['grads_accumulators'] = {k: tolist(v) for k, v in model.get_tf_params(scope='grads_accumulators').items()}
    rec['log'] = {'train': log['train'], 'val': log['val'], 'feg': [float(v) for v in log['feg']]}
    rec['epoch_'], rec['iter_'] = int(model.epoch_), int(model.iter_)
    return rec


def main():
    tf1shim.default_random_provider = provider
    tf1shim.install()
    sys.path[:0] = ['/root/reference', '/root/reference/boltzmann_machines']
    import boltzmann_machines as ref                     # the reference package, unmodified
    assert ref.__file__.startswith('/root/reference/'), ref.__file__
    work = tempfile.mkdtemp(prefix='bm_golden_')
    cwd = os.getcwd()
    try:
        os.chdir(work)
        recs = [run_case(ref.rbm, c, work) for c in cases()]
    finally:
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)
    out = {'source': 'yell/boltzmann-machines (boltzmann_machines/rbm, layers.py, base/tf_model.py) executed unmodified on '
                     'oracle/tf1shim.py; random draws from the shared Philox layout (tests/golden/make_reference_rbm_golden.py)',
           'cases': recs}
    path = os.path.join(HERE, 'reference_rbm_cases.json')
    with open(path, 'w') as fh:
        json.dump(out, fh)
    print('wrote', path, os.path.getsize(path), 'bytes;', len(recs), 'cases')
    for r in recs:
        print(' ', r['name'], 'iter_', r['iter_'], 'train log', r['log']['train'][-1:] , 'feg', r['log']['feg'])


if __name__ == '__main__':
    main()
