"""Deep Boltzmann Machine: host-side training loop over the native engine.

Interface parity target: /root/reference/boltzmann_machines/dbm.py (constructor kwargs
:89-99, ``load_rbms`` :207-231, composition of the initial weights from pre-trained RBMs
:266-291, epoch/batch loop :793-857, public methods :859-957).  What the reference runs
inside ``session.run`` -- the mean-field E-step, the PCD particle update, gradients,
sparsity, momentum, max-norm, AIS, the variational bound -- happens in the engine
(``bm_dbm_*`` in include/bm.h).

Session semantics: in the reference every public call opens a session, restores the
variables from disk and (unless it saves) throws its changes away.  ``transform``,
``reconstruct``, ``log_proba``, ``log_Z`` and ``sample_v(save_model=False)`` therefore must
not leave traces in the persistent state (variational parameters, particles): they snapshot
and restore it around the call.
"""
import numpy as np

from .base import run_in_tf_session, get_engine_factory
from .ebm import EnergyBasedModel
from .layers import BernoulliLayer
from .utils import (make_list_from, write_during_training, batch_bounds, epoch_iter,
                    log_sum_exp, log_diff_exp, log_mean_exp, log_std_exp)
from .utils.utils import _maybe_bar


def _sfx(i):
    """TF uniquifies repeated variable names: W, W_1, W_2, ... (dbm_mnist.py:367-371)."""
    return '' if i == 0 else '_{0}'.format(i)


class DBM(EnergyBasedModel):
    """Deep Boltzmann Machine trained with PCD and mean-field variational inference.

    Parameters (identical names and meaning to the reference)
    ----------
    rbms : list of pre-trained RBMs, visible side first
    n_particles : number of persistent Markov chains
    v_particle_init, h_particles_init : optional initial particle states
    n_gibbs_steps : PCD sweeps per update (scalar or per-epoch list)
    max_mf_updates, mf_tol : mean-field iteration cap and sup-norm tolerance
    learning_rate, momentum : scalar or per-epoch list
    max_epoch, batch_size : ints (``len(X)`` should be divisible by ``batch_size``)
    l2, max_norm : weight decay; column-norm constraint
    sample_v_states, sample_h_states : bool, list of bool per hidden layer
    sparsity_target, sparsity_cost : scalar or per-layer; sparsity_damping
    train_metrics_every_iter, val_metrics_every_epoch, verbose, save_after_each_epoch
    display_filters, display_particles, v_shape : accepted for compatibility
    """
    _MUTABLE_PREFIXES = ('mu', 'v', 'h')

    def __init__(self, rbms=None,
                 n_particles=100, v_particle_init=None, h_particles_init=None,
                 n_gibbs_steps=5, max_mf_updates=10, mf_tol=1e-7,
                 learning_rate=0.0005, momentum=0.9, max_epoch=10, batch_size=100,
                 l2=0., max_norm=np.inf,
                 sample_v_states=True, sample_h_states=None,
                 sparsity_target=0.1, sparsity_cost=0., sparsity_damping=0.9,
                 train_metrics_every_iter=10, val_metrics_every_epoch=1,
                 verbose=False, save_after_each_epoch=True,
                 display_filters=0, display_particles=0, v_shape=(28, 28),
                 model_path='dbm_model/', *args, **kwargs):
        super(DBM, self).__init__(model_path=model_path, *args, **kwargs)
        self.n_layers_ = len(rbms) if rbms is not None else None
        self.n_visible_ = None
        self.n_hiddens_ = []
        self.layers_ = None                  # unit kinds/params, persisted so that a loaded model can run
        self._rbms = None
        self.load_rbms(rbms)

        self.n_particles = n_particles
        self._v_particle_init = v_particle_init
        self._h_particles_init = h_particles_init

        self.n_gibbs_steps = make_list_from(n_gibbs_steps)
        self.max_mf_updates = max_mf_updates
        self.mf_tol = mf_tol

        self.learning_rate = make_list_from(learning_rate)
        self.momentum = make_list_from(momentum)
        self.max_epoch = max_epoch
        self.batch_size = batch_size
        self.l2 = l2
        self.max_norm = max_norm

        self.sample_v_states = sample_v_states
        self.sample_h_states = sample_h_states or ([True] * self.n_layers_ if self.n_layers_ else None)

        self.sparsity_target = make_list_from(sparsity_target)
        self.sparsity_cost = make_list_from(sparsity_cost)
        if self.n_layers_ is not None and self.n_layers_ > 1:
            for x in (self.sparsity_target, self.sparsity_cost):
                if len(x) == 1:
                    x *= self.n_layers_
        self.sparsity_damping = sparsity_damping

        self.train_metrics_every_iter = train_metrics_every_iter
        self.val_metrics_every_epoch = val_metrics_every_epoch
        self.verbose = verbose
        self.save_after_each_epoch = save_after_each_epoch

        for nh in self.n_hiddens_:
            assert nh >= display_filters
        self.display_filters = display_filters
        assert display_particles <= self.n_particles
        self.display_particles = display_particles
        self.v_shape = tuple(v_shape)
        if len(self.v_shape) == 2:
            self.v_shape = self.v_shape + (1,)

        self.epoch_ = 0
        self.iter_ = 0
        self.n_samples_generated_ = 0

    # ---- composition from pre-trained RBMs (dbm.py:207-231, 266-291) ---------------------
    def load_rbms(self, rbms):
        if rbms is None:
            return
        self._rbms = rbms
        self.n_layers_ = len(rbms)
        self.n_visible_ = rbms[0].n_visible
        self.n_hiddens_ = [rbm.n_hidden for rbm in rbms]
        self._W_init, self._vb_init, self._hb_init = [], [], []
        for rbm in rbms:
            w = rbm.get_tf_params(scope='weights')
            self._W_init.append(w['W']); self._vb_init.append(w['vb']); self._hb_init.append(w['hb'])
        self._v_layer = rbms[0]._v_layer
        self._h_layers = [rbm._h_layer for rbm in rbms]
        self._v_layer.dtype = self.dtype
        for h in self._h_layers:
            h.dtype = self.dtype
        desc = lambda layer: dict(kind=layer.kind, **{k: (np.asarray(v).tolist() if hasattr(v, '__iter__') else v)
                                                      for k, v in layer.engine_params().items()})
        self.layers_ = dict(v=desc(self._v_layer), h=[desc(h) for h in self._h_layers])
        if getattr(self, 'sample_h_states', None) is None:
            self.sample_h_states = [True] * self.n_layers_

    def _composed_init(self):
        """Halve what is counted twice when RBMs are stacked (dbm.py:266-291)."""
        L = self.n_layers_
        W_init, hb_init = [], []
        vb_init = np.array(self._vb_init[0], dtype=self._np_dtype)
        for i in range(L):
            W = np.array(self._W_init[i], dtype=self._np_dtype)
            vb = np.array(self._vb_init[i], dtype=self._np_dtype)
            hb = np.array(self._hb_init[i], dtype=self._np_dtype)
            if 0 < i < L - 1:                  # intermediate RBMs: both directions were doubled
                W *= 0.5; vb *= 0.5; hb *= 0.5
            W_init.append(W)
            if i == 0:
                hb_init.append(0.5 * hb)
            else:
                hb_init[i - 1] = hb_init[i - 1] + 0.5 * vb
                hb_init.append(0.5 * hb if i < L - 1 else hb)
        return vb_init, W_init, hb_init

    # ---- engine wiring ---------------------------------------------------------------------------
    def _scopes(self):
        L = self.n_layers_ or 0
        per = lambda stem: tuple(stem + _sfx(i) for i in range(L))
        return {
            'weights': ('vb',) + per('W') + per('hb'),
            'grads_accumulators': ('dvb',) + per('dW') + per('dhb'),
            'variational_params': per('mu'),
            'hidden_means_accumulators': per('q_means') + per('mu_means'),
            'negative_particles': ('v',) + per('h'),
        }

    def _tf_name(self, scope, name):
        # the hidden particles are created inside `tf.name_scope('h_particle')` (dbm.py:371-383): TF names them
        # negative_particles/h_particle/h, negative_particles/h_particle_1/h, ...
        if scope == 'negative_particles' and name.startswith('h'):
            i = name[2:]
            return 'h_particle{0}/h'.format('_' + i if i else '')
        return name

    def _engine_cfg(self):
        if self.layers_ is None:
            raise RuntimeError('the DBM has no layer description: call `load_rbms` (or load a saved model)')
        lay = self.layers_
        cfg = dict(
            n_visible=int(self.n_visible_), n_hiddens=[int(h) for h in self.n_hiddens_],
            v_kind=lay['v']['kind'], h_kinds=[h['kind'] for h in lay['h']],
            h_n_samples=[float(h.get('n_samples', 100.)) for h in lay['h']],
            dtype=self.dtype, n_particles=int(self.n_particles), batch_size=int(self.batch_size),
            max_mf_updates=int(self.max_mf_updates), mf_tol=float(self.mf_tol),
            l2=float(self.l2), max_norm=float(self.max_norm),
            sample_v=bool(self.sample_v_states), sample_h=[bool(s) for s in self.sample_h_states],
            sparsity_target=[float(x) for x in self.sparsity_target],
            sparsity_cost=[float(x) for x in self.sparsity_cost],
            sparsity_damping=float(self.sparsity_damping),
        )
        if 'sigma' in lay['v']:
            cfg['sigma'] = np.broadcast_to(np.asarray(lay['v']['sigma'], dtype=np.float64), (self.n_visible_,)).copy()
        return cfg

    def _make_engine(self):
        return get_engine_factory('dbm')(self._engine_cfg())

    _make_tf_model = _make_engine

    def _init_engine_vars(self):
        if self._rbms is None:
            raise RuntimeError('`load_rbms` must be called before the first `fit`/`init`')
        dt = self._np_dtype
        vb, Ws, hbs = self._composed_init()
        d = {'vb': vb}
        for i in range(self.n_layers_):
            d['W' + _sfx(i)] = Ws[i]
            d['hb' + _sfx(i)] = hbs[i]
        self._engine.set_params(d)
        # persistent particles: the layers' own initialisers unless given (dbm.py:362-383)
        self._engine.init_particles(self.make_random_seed() if self.random_seed is not None
                                    else int(np.random.SeedSequence().generate_state(1)[0]))
        given = {}
        if self._v_particle_init is not None:
            given['v'] = np.asarray(self._v_particle_init, dtype=dt)
        if self._h_particles_init is not None:
            for i, h in enumerate(self._h_particles_init):
                if h is not None:
                    given['h' + _sfx(i)] = np.asarray(h, dtype=dt).reshape(self.n_particles, self.n_hiddens_[i])
        if given:
            self._engine.set_params(given)

    def _snapshot(self):
        names = [n for sc in ('variational_params', 'negative_particles') for n in self._scopes()[sc]]
        return self._engine.get_params(names)

    # ---- schedules -----------------------------------------------------------------------------------
    def _scheduled(self, values):
        return values[min(self.epoch_, len(values) - 1)]

    def _step_args(self, n_gibbs_steps=None):
        k = self._scheduled(self.n_gibbs_steps) if n_gibbs_steps is None else n_gibbs_steps
        return dict(lr=float(self._scheduled(self.learning_rate)), momentum=float(self._scheduled(self.momentum)),
                    k=int(k), seed=self._call_seed)

    # ---- training ------------------------------------------------------------------------------------
    def _train_epoch(self, X):
        msres, n_mfs = [], []
        for lo, hi in _maybe_bar(batch_bounds(len(X), self.batch_size), self.verbose, leave=False, ncols=64, desc='epoch'):
            self.iter_ += 1
            report = ('msre', 'n_mf_updates') if (self.train_metrics_every_iter and
                                                  self.iter_ % self.train_metrics_every_iter == 0) else ()
            got = self._engine.train_step(X[lo:hi], tick=self._next_tick(), metrics=report, **self._step_args())
            if report:
                msres.append(got['msre']); n_mfs.append(got['n_mf_updates'])
                self._log_scalars('train', self.iter_, {'mean_squared_recon_error': got['msre'],        # dbm.py:636-639
                                                        'n_mf_updates': got['n_mf_updates']})
        return (float(np.mean(msres)) if msres else None, float(np.mean(n_mfs)) if n_mfs else None)

    def _run_val_metrics(self, X_val):
        a = self._step_args()
        msres, n_mfs = [], []
        for lo, hi in batch_bounds(len(X_val), self.batch_size):
            got = self._engine.val_metrics(X_val[lo:hi], k=a['k'], seed=a['seed'], tick=self._next_tick())
            msres.append(got['msre']); n_mfs.append(got['n_mf_updates'])
        self._log_scalars('val', self.iter_, {'mean_squared_recon_error': np.mean(msres),              # dbm.py:818-823
                                              'n_mf_updates': np.mean(n_mfs)})
        return float(np.mean(msres)), float(np.mean(n_mfs))

    def _fit(self, X, X_val=None, *args, **kwargs):
        X = np.ascontiguousarray(X, dtype=self._np_dtype)
        if X_val is not None:
            X_val = np.ascontiguousarray(X_val, dtype=self._np_dtype)
        val_msre, val_n_mf = None, None
        for self.epoch_ in epoch_iter(start_epoch=self.epoch_, max_epoch=self.max_epoch, verbose=self.verbose):
            train_msre, train_n_mf = self._train_epoch(X)
            if X_val is not None and self.epoch_ % self.val_metrics_every_epoch == 0:
                val_msre, val_n_mf = self._run_val_metrics(X_val)
            if self.verbose:
                s = 'epoch: {0:{1}}/{2}'.format(self.epoch_, len(str(self.max_epoch)), self.max_epoch)
                if train_msre:
                    s += '; msre: {0:.5f}'.format(train_msre)
                if train_n_mf:
                    s += '; n_mf_upds: {0:.1f}'.format(train_n_mf)
                if val_msre:
                    s += '; val.msre: {0:.5f}'.format(val_msre)
                if val_n_mf:
                    s += '; val.n_mf_upds: {0:.1f}'.format(val_n_mf)
                write_during_training(s)
            if self.save_after_each_epoch:
                self._save_model(global_step=self.epoch_)

    # ---- queries -------------------------------------------------------------------------------------
    def _per_batch(self, X, width, fn, desc, dtype=None):
        X = np.ascontiguousarray(X, dtype=self._np_dtype)
        out = np.zeros((len(X), width) if width else (len(X),), dtype=dtype or self._np_dtype)
        keep = self._snapshot()
        try:
            for lo, hi in _maybe_bar(batch_bounds(len(X), self.batch_size), self.verbose, leave=False, ncols=64, desc=desc):
                out[lo:hi] = fn(X[lo:hi])
        finally:
            self._engine.set_params(keep)
        return out

    @run_in_tf_session()
    def transform(self, X, np_dtype=None):
        """Activation probabilities of the last hidden layer after the mean-field E-step."""
        return self._per_batch(X, self.n_hiddens_[-1], self._engine.transform, 'transform', np_dtype)

    @run_in_tf_session(update_seed=True)
    def reconstruct(self, X):
        """p(v | h_0 = q) with q the mean-field posterior of the first hidden layer."""
        return self._per_batch(X, self.n_visible_, self._engine.reconstruct, 'reconstruction')

    @run_in_tf_session(update_seed=True)
    def sample_v(self, n_gibbs_steps=0, save_model=False):
        """Visible activation probabilities of the particles after ``n_gibbs_steps`` sweeps."""
        keep = None if save_model else self._snapshot()
        v = self._engine.sample_v(int(n_gibbs_steps), self._call_seed, self._next_tick())
        if save_model:
            self.n_samples_generated_ += n_gibbs_steps
            self._save_model()
        else:
            self._engine.set_params(keep)
        return v

    @run_in_tf_session(update_seed=True)
    def log_Z(self, n_betas=100, n_runs=100, n_gibbs_steps=5):
        """Annealed-importance-sampling estimate of the log partition function (2 binary hidden
        layers; state space h_1 with v and h_2 summed out analytically).

        Returns ``log_mean, (log_low, log_high), values``: log of the mean estimate, of the mean
        minus / plus one standard deviation, and the ``n_runs`` individual estimates.
        """
        assert self.n_layers_ == 2
        assert self.layers_['v']['kind'] == 'bernoulli' and all(h['kind'] == 'bernoulli' for h in self.layers_['h'])
        ais_seed = self.make_random_seed()           # dbm.py:701
        values = self._engine.ais(int(n_runs), int(n_betas), int(n_gibbs_steps), ais_seed)
        log_mean = log_mean_exp(values)
        log_std = log_std_exp(values, log_mean_exp_x=log_mean)
        log_high = log_sum_exp([log_std, log_mean])
        log_low = log_diff_exp([log_std, log_mean])[0]
        return log_mean, (log_low, log_high), values

    @run_in_tf_session()
    def log_proba(self, X_test, log_Z):
        """Variational lower bound on log p(v) for the rows of ``X_test`` given an estimate of log Z."""
        assert self.n_layers_ == 2
        assert self.layers_['v']['kind'] == 'bernoulli' and all(h['kind'] == 'bernoulli' for h in self.layers_['h'])
        P = self._per_batch(X_test, 0, self._engine.log_proba, 'log_proba', np.float64)
        return P - log_Z
