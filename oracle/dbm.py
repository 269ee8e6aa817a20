"""CPU restatement (numpy) of the reference's DBM path -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/boltzmann_machines/dbm.py: composition from RBMs happens in
the package's host code (dbm.py:266-291 is pure numpy); this engine holds the TF
variables (weights, momentum accumulators, variational parameters mu, persistent
particles) and performs what the reference runs inside ``session.run``:

  * layer-wise Gibbs step                      dbm.py:385-427
  * mean-field E-step (+ its quirks)           dbm.py:429-478
  * PCD particle update                        dbm.py:480-509
  * gradients, sparsity, momentum, max-norm    dbm.py:511-513, 550-615
  * msre / reconstruction                      dbm.py:625-633
  * sample_v                                   dbm.py:641-648
  * AIS                                        dbm.py:650-736
  * variational lower bound                    dbm.py:738-759

PARITY STATUS: pinned against the reference's OWN SOURCE.  tests/golden/reference_dbm_cases.json holds the results of
yell/boltzmann-machines' unmodified `DBM` class -- greedy pre-training of two RBMs, `fit` (mean-field E-step with its
stale-mu start, PCD particles with a variable number of Gibbs steps, the sparsity update, max-norm, momentum),
`transform`, `reconstruct`, `sample_v`, `log_proba` and `log_Z` (AIS) -- executed with oracle/tf1shim.py in place of
TensorFlow 1.3 and the random ops answered from the Philox layout of oracle/philox.py;
tests/test_z_reference_golden.py replays the scenario through the host mirror on this oracle (and, on the GPU, on the CUDA
engine): every variable of every scope after `fit`, the MSRE / n_mf_updates logs, the query results and the AIS
log-weights (to 5e-6) agree.  The quirks the restatement carries over are therefore confirmed by the reference's code:
the first mean-field sweep reads the *previous batch's* mu (dbm.py:459-467), the sparsity update indexes element i of
the per-unit sum vector (dbm.py:581-586), X W_0 is doubled in the approximate-inference pass even for a single layer
(dbm.py:438), validation metrics advance the persistent chains (dbm.py:523).  NOT pinned: TensorFlow's own draw order
(the engine defines its own counter layout) and TensorFlow's kernels.  AIS importance weights are accumulated in
float64 here (float32 in the reference's graph).
"""
import numpy as np

from . import philox as P
from .rbm import sigmoid, softplus, bf16_round, multinomial_counts


class OracleDBM(object):
    def __init__(self, cfg):
        self.cfg = c = dict(cfg)
        self.V = int(c['n_visible'])
        self.Hs = [int(h) for h in c['n_hiddens']]
        self.L = len(self.Hs)
        self.dt = np.dtype(c.get('dtype', 'float32'))
        self.v_kind = c.get('v_kind', 'bernoulli')
        self.h_kinds = list(c.get('h_kinds', ['bernoulli'] * self.L))
        self.h_n_samples = list(c.get('h_n_samples', [100.] * self.L))
        self.M = int(c['n_particles'])
        self.B = int(c['batch_size'])
        sig = c.get('sigma', None)
        self.sigma = None if sig is None else np.broadcast_to(np.asarray(sig, dtype=self.dt), (self.V,)).copy()
        z = lambda *s: np.zeros(s, dtype=self.dt)
        sizes = [self.V] + self.Hs
        self.p = {'vb': z(self.V), 'dvb': z(self.V)}
        for i in range(self.L):
            s = self._sfx(i)
            self.p['W' + s] = z(sizes[i], sizes[i + 1]); self.p['dW' + s] = z(sizes[i], sizes[i + 1])
            self.p['hb' + s] = z(self.Hs[i]); self.p['dhb' + s] = z(self.Hs[i])
            self.p['mu' + s] = z(self.B, self.Hs[i])
            self.p['q_means' + s] = z(self.Hs[i]); self.p['mu_means' + s] = z(self.Hs[i])
            self.p['h' + s] = z(self.M, self.Hs[i])
        self.p['v'] = z(self.M, self.V)

    @staticmethod
    def _sfx(i):
        """TF uniquifies repeated variable names: W, W_1, W_2, ... (dbm_mnist.py:367-371)."""
        return '' if i == 0 else '_{0}'.format(i)

    # -- variables -------------------------------------------------------------
    def set_params(self, d):
        for k, v in d.items():
            if k == 'sigma':
                continue
            self.p[k] = np.array(v, dtype=self.dt).reshape(self.p[k].shape)

    def get_params(self, names=None):
        names = list(self.p) if names is None else list(names)
        return {k: self.p[k].copy() for k in names}

    # -- data parallelism (the engine's contract, csrc/bm_dbm.cu; the reference is single-device) ----------------
    def set_shard(self, rank, world, allreduce_sum, allreduce_max):
        """This oracle holds shard `rank` of `world`: batch_size rows of every batch / of mu and n_particles particles
        (the global model has `world` times as many).  Sampling sites are keyed by the global particle index, the
        mean-field test takes the max over the shards, the statistics of the update are summed over them."""
        self.shard = (int(rank), int(world), allreduce_sum, allreduce_max)

    def _particle_row0(self):
        sh = getattr(self, 'shard', None)
        return 0 if sh is None else sh[0] * self.M

    def init_particles(self, seed):
        """layer.init(batch_size=n_particles) for every layer (dbm.py:362-383; layers.py:43-45,
        59-63,78-82): Bernoulli -> U[0,1), Multinomial -> U[0,1)/sum, Gaussian -> sigma*N(0,1)."""
        kinds = [self.v_kind] + self.h_kinds
        sizes = [self.V] + self.Hs
        names = ['v'] + ['h' + self._sfx(i) for i in range(self.L)]
        r0 = self._particle_row0()
        for idx, (kind, n, name) in enumerate(zip(kinds, sizes, names)):
            if kind == 'gaussian':
                t = P.normal_at(self.M, n, seed, P.SITE_PARTICLE_INIT, idx, 0, r0) * self.sigma[None, :]
            else:
                t = P.uniform_at(self.M, n, seed, P.SITE_PARTICLE_INIT, idx, 0, r0)
                if kind == 'multinomial':
                    t = t / t.sum(dtype=np.float64)
            self.p[name] = t.astype(self.dt)

    def W(self, i): return self.p['W' + self._sfx(i)]
    def hb(self, i): return self.p['hb' + self._sfx(i)]

    # -- layers ------------------------------------------------------------------
    def _act(self, kind, x, b, n_samples=100., beta=None):
        if beta is not None:          # AIS: activation(beta * x, beta * b)
            x, b = beta * x, beta * b
        if kind == 'bernoulli':
            return sigmoid((x + b).astype(self.dt))
        if kind == 'multinomial':
            zc = x + b
            zc = zc - zc.max(axis=1, keepdims=True)
            e = np.exp(zc)
            return (self.dt.type(n_samples) * e / e.sum(axis=1, keepdims=True)).astype(self.dt)
        return (x * self.sigma + b).astype(self.dt)

    def _sample(self, kind, means, n_samples, seed, site, t, tick):
        rows, n = means.shape
        if kind == 'bernoulli':
            u = P.uniform_at(rows, n, seed, site, t, tick, getattr(self, '_row0', 0))
            return (u.astype(self.dt) < means).astype(self.dt)
        if kind == 'multinomial':
            probs = (means / means.sum(axis=1, keepdims=True)).astype(np.float32)
            return multinomial_counts(probs, int(n_samples), seed, site, t, tick).astype(self.dt)
        eps = P.normal_at(rows, n, seed, site, t, tick, getattr(self, '_row0', 0)).astype(self.dt)
        return (means + self.sigma * eps).astype(self.dt)

    def gibbs_step(self, v, H, update_v, sample, seed, t, tick):
        """dbm.py:385-427.  Returns (v_new, H_new)."""
        L, sh = self.L, self.cfg.get('sample_h', [True] * self.L)
        Hn = [None] * L
        T = v @ self.W(0)
        if L >= 2:
            T = T + H[1] @ self.W(1).T
        Hn[0] = self._act(self.h_kinds[0], T, self.hb(0), self.h_n_samples[0])
        if sample and sh[0]:
            Hn[0] = self._sample(self.h_kinds[0], Hn[0], self.h_n_samples[0], seed, P.SITE_DBM_H + 0, t, tick)
        for i in range(1, L - 1):
            T = Hn[i - 1] @ self.W(i) + H[i + 1] @ self.W(i + 1).T
            Hn[i] = self._act(self.h_kinds[i], T, self.hb(i), self.h_n_samples[i])
            if sample and sh[i]:
                Hn[i] = self._sample(self.h_kinds[i], Hn[i], self.h_n_samples[i], seed, P.SITE_DBM_H + i, t, tick)
        if L >= 2:
            T = Hn[L - 2] @ self.W(L - 1)
            Hn[L - 1] = self._act(self.h_kinds[L - 1], T, self.hb(L - 1), self.h_n_samples[L - 1])
            if sample and sh[L - 1]:
                Hn[L - 1] = self._sample(self.h_kinds[L - 1], Hn[L - 1], self.h_n_samples[L - 1], seed,
                                         P.SITE_DBM_H + L - 1, t, tick)
        v_new = v
        if update_v:
            v_new = self._act(self.v_kind, Hn[0] @ self.W(0).T, self.p['vb'])
            if sample and self.cfg.get('sample_v', True):
                v_new = self._sample(self.v_kind, v_new, 100., seed, P.SITE_DBM_V, t, tick)
        return v_new, Hn

    # -- E-step --------------------------------------------------------------------
    def mean_field(self, X):
        """dbm.py:429-478.  Updates self mu; returns the number of sweeps."""
        L = self.L
        rows = X.shape[0]
        mu = [self.p['mu' + self._sfx(i)][:rows] for i in range(L)]         # previous batch's values
        mu_new = []
        T = None
        for i in range(L):                                                  # :434-446
            if i == 0:
                T = self.dt.type(2.) * (X @ self.W(0))
            else:
                T = T @ self.W(i)
                if i < L - 1:
                    T = T * self.dt.type(2.)
            T = self._act(self.h_kinds[i], T, self.hb(i), self.h_n_samples[i])
            mu_new.append(T)
        step = 0
        tol = self.dt.type(self.cfg.get('mf_tol', 1e-7))
        def spread(mu, mu_new):
            d = max(np.max(np.abs(a - b)) for a, b in zip(mu, mu_new))
            sh = getattr(self, 'shard', None)
            return d if sh is None else self.dt.type(sh[3](np.array([d], dtype=np.float64))[0])
        while step < int(self.cfg.get('max_mf_updates', 10)) and spread(mu, mu_new) > tol:      # :449-452
            _, Hn = self.gibbs_step(X, mu, update_v=False, sample=False, seed=0, t=0, tick=0)
            mu, mu_new = Hn, mu                                                    # :455-457 (swap)
            step += 1
        for i in range(L):
            self.p['mu' + self._sfx(i)][:rows] = mu[i]
        return step

    # -- M-step ---------------------------------------------------------------------
    def particles_update(self, n_steps, sample, seed, tick, t0=0, commit=True):
        """dbm.py:480-509.  Returns (v, H) after n_steps; commits them to the persistent particles."""
        v = self.p['v']
        H = [self.p['h' + self._sfx(i)] for i in range(self.L)]
        self._row0 = self._particle_row0()
        try:
            for s in range(int(n_steps)):
                v, H = self.gibbs_step(v, H, update_v=True, sample=sample, seed=seed, t=t0 + s + 1, tick=tick)
        finally:
            self._row0 = 0
        if commit:
            self.p['v'] = v
            for i in range(self.L):
                self.p['h' + self._sfx(i)] = H[i]
        return v, H

    def _max_norm(self, T):
        n = np.sqrt(np.sum(np.square(T), axis=0))                                   # :511-513
        mn = self.dt.type(self.cfg.get('max_norm', np.inf))
        return (T * np.minimum(n, mn) / np.maximum(n, self.dt.type(1e-8))).astype(self.dt)

    def reconstruction(self, rows):
        mu0 = self.p['mu'][:rows]
        return self._act(self.v_kind, mu0 @ self.W(0).T, self.p['vb'])              # :626-628

    def train_step(self, X, lr, momentum, k, seed, tick, metrics=()):
        """One ``session.run(train_op)`` (dbm.py:515-622)."""
        c, p, dt, L = self.cfg, self.p, self.dt, self.L
        X = np.asarray(X, dtype=dt)
        rows = X.shape[0]
        n_mf = self.mean_field(X)
        self.particles_update(k, True, seed, tick)
        out = None
        if metrics:
            out = {}
            if 'msre' in metrics:
                out['msre'] = float(np.mean(np.square(X - self.reconstruction(rows)), dtype=np.float64))
            if 'n_mf_updates' in metrics:
                out['n_mf_updates'] = float(n_mf)
        N, Mp = dt.type(self.B), dt.type(self.M)          # the reference divides by the *configured* sizes (:254-255)
        mu = [p['mu' + self._sfx(i)][:rows] for i in range(L)]
        Hp = [p['h' + self._sfx(i)] for i in range(L)]
        l2 = dt.type(c.get('l2', 0.))
        shard = getattr(self, 'shard', None)
        if shard is None:
            dvb = X.mean(axis=0) - p['v'].mean(axis=0)                                   # :553
            dW = [(X.T @ mu[0]) / N - (p['v'].T @ Hp[0]) / Mp - l2 * self.W(0)]          # :558-560
            for i in range(1, L):
                dW.append((mu[i - 1].T @ mu[i]) / N - (Hp[i - 1].T @ Hp[i]) / Mp - l2 * self.W(i))   # :566-568
            dhb = [mu[i].mean(axis=0) - Hp[i].mean(axis=0) for i in range(L)]            # :575
            q_sums, mu_sums = [Hp[i].sum(axis=0) for i in range(L)], [mu[i].sum(axis=0) for i in range(L)]
        else:
            # the same quantities from sums over the shards: global sizes, one sum-allreduce (csrc/bm_dbm.cu)
            world, allsum = shard[1], shard[2]
            N, Mp, rows_g = dt.type(self.B * world), dt.type(self.M * world), dt.type(rows * world)
            pos, neg = [X] + mu[:-1], [p['v']] + Hp[:-1]
            parts = [(pos[i].T @ mu[i]) / N - (neg[i].T @ Hp[i]) / Mp for i in range(L)]
            parts += [mu[i].sum(axis=0) for i in range(L)] + [Hp[i].sum(axis=0) for i in range(L)]
            parts += [X.sum(axis=0), p['v'].sum(axis=0)]
            flat = allsum(np.concatenate([np.ravel(a) for a in parts]).astype(dt))
            out_parts, off = [], 0
            for a in parts:
                out_parts.append(flat[off:off + a.size].reshape(a.shape)); off += a.size
            G, mu_sums, q_sums = out_parts[:L], out_parts[L:2 * L], out_parts[2 * L:3 * L]
            x_sum, v_sum = out_parts[3 * L], out_parts[3 * L + 1]
            dvb = x_sum / rows_g - v_sum / Mp
            dW = [G[i] - l2 * self.W(i) for i in range(L)]
            dhb = [mu_sums[i] / rows_g - q_sums[i] / Mp for i in range(L)]
            if out is not None and 'msre' in out:
                out['msre'] = float(allsum(np.array([out['msre']], dtype=np.float64))[0] / world)
        damp = dt.type(c.get('sparsity_damping', 0.9))
        targets = c.get('sparsity_target', [0.1] * L)
        costs = c.get('sparsity_cost', [0.] * L)
        for i in range(L):                                                           # :580-590 (element i: sic)
            s = self._sfx(i)
            qv = q_sums[i]
            p['q_means' + s] = (damp * p['q_means' + s] + (dt.type(1) - damp) * qv[i]).astype(dt)
            mv = mu_sums[i]
            p['mu_means' + s] = (damp * p['mu_means' + s] + (dt.type(1) - damp) * mv[i]).astype(dt)
            pen = dt.type(costs[i]) * (p['q_means' + s] - dt.type(targets[i]))
            pen = pen + dt.type(costs[i]) * (p['mu_means' + s] - dt.type(targets[i]))
            dW[i] = dW[i] - pen[None, :]
            dhb[i] = dhb[i] - pen
        lr, mom = dt.type(lr), dt.type(momentum)
        p['dvb'] = (lr * (mom * p['dvb'] + dvb)).astype(dt)                          # :595-596
        p['vb'] = (p['vb'] + p['dvb']).astype(dt)
        for i in range(L):
            s = self._sfx(i)
            p['dW' + s] = (lr * (mom * p['dW' + s] + dW[i])).astype(dt)              # :602-606
            p['W' + s] = self._max_norm(p['W' + s] + p['dW' + s])
            p['dhb' + s] = (lr * (mom * p['dhb' + s] + dhb[i])).astype(dt)           # :613-614
            p['hb' + s] = (p['hb' + s] + p['dhb' + s]).astype(dt)
        return out

    # -- read-only queries (the caller restores mu / particles afterwards) -------------------------
    def transform(self, X):
        X = np.asarray(X, dtype=self.dt)
        self.mean_field(X)
        return self.p['mu' + self._sfx(self.L - 1)][:X.shape[0]].copy()              # :526-528

    def reconstruct(self, X):
        X = np.asarray(X, dtype=self.dt)
        self.mean_field(X)
        return self.reconstruction(X.shape[0])

    def val_metrics(self, X, k, seed, tick):
        """dbm.py:810-816: msre / n_mf_updates sit under control_dependencies on the mean-field AND the
        particle updates (dbm.py:523), so evaluating them on validation data also advances the
        persistent chains -- inside ``fit`` that state is kept."""
        X = np.asarray(X, dtype=self.dt)
        n_mf = self.mean_field(X)
        self.particles_update(k, True, seed, tick)
        msre = float(np.mean(np.square(X - self.reconstruction(X.shape[0])), dtype=np.float64))
        return {'msre': msre, 'n_mf_updates': float(n_mf)}

    def sample_v(self, k, seed, tick):
        """dbm.py:641-648: k sampled sweeps, then k more without sampling; v <- those means."""
        self.particles_update(k, True, seed, tick)
        v, _ = self.particles_update(k, False, seed, tick, t0=int(k), commit=False)
        self.p['v'] = v
        return v.copy()

    def log_proba(self, X):
        """dbm.py:738-759 (without the -log_Z the caller subtracts)."""
        X = np.asarray(X, dtype=self.dt)
        rows = X.shape[0]
        self.mean_field(X)
        mu0, mu1 = self.p['mu'][:rows], self.p['mu_1'][:rows]
        minus_E = ((X @ self.W(0)) * mu0).sum(axis=1) + ((mu0 @ self.W(1)) * mu1).sum(axis=1)
        minus_E = minus_E + X @ self.p['vb'] + mu0 @ self.hb(0) + mu1 @ self.hb(1)
        ent = 0.
        for m in (mu0, mu1):
            s = np.clip(m, 1e-7, 1. - 1e-7)
            ent = ent + (-s * np.log(s) - (1. - s) * np.log(1. - s)).sum(axis=1)
        return (minus_E + ent).astype(np.float64)

    # -- AIS (dbm.py:650-736) ---------------------------------------------------------------------
    def _log_p_star(self, x, beta):
        T1 = (x @ self.hb(0)).astype(np.float64) * beta
        T2 = softplus((beta * (x @ self.W(0).T + self.p['vb'])).astype(np.float64)).sum(axis=1)
        T3 = softplus((beta * (x @ self.W(1) + self.hb(1))).astype(np.float64)).sum(axis=1)
        return T1 + T2 + T3

    def _ais_transition(self, x, beta, k, seed, it):
        sh = self.cfg.get('sample_h', [True] * self.L)
        for s in range(int(k)):
            tick = it * int(k) + s
            v = self._act(self.v_kind, x @ self.W(0).T, self.p['vb'], beta=beta)
            if self.cfg.get('sample_v', True):
                v = self._sample(self.v_kind, v, 100., seed, P.SITE_AIS_V, 0, tick)
            h2 = self._act(self.h_kinds[1], x @ self.W(1), self.hb(1), beta=beta)
            if sh[1]:
                h2 = self._sample(self.h_kinds[1], h2, 100., seed, P.SITE_AIS_H2, 0, tick)
            x = self._act(self.h_kinds[0], v @ self.W(0) + h2 @ self.W(1).T, self.hb(0), beta=beta)
            if sh[0]:
                x = self._sample(self.h_kinds[0], x, 100., seed, P.SITE_AIS_H1, 0, tick)
        return x

    def ais(self, n_runs, n_betas, k, seed, first_run=0):
        """``first_run``: these are runs [first_run, first_run + n_runs) of a longer ladder (the runs are independent
        chains; run r draws from row r of the AIS sites), which is how the engine shards them over GPUs."""
        assert self.L == 2
        self._row0 = int(first_run)
        try:
            return self._ais(n_runs, n_betas, k, seed)
        finally:
            self._row0 = 0

    def _ais(self, n_runs, n_betas, k, seed):
        dt = self.dt
        delta = dt.type(1. / n_betas)
        u = P.uniform_at(int(n_runs), self.Hs[0], seed, P.SITE_AIS_INIT, 0, 0, self._row0)
        x = (u < np.float32(0.5)).astype(dt)                                        # :700-702
        it = 0
        x = self._ais_transition(x, delta, k, seed, it); it += 1                     # :705
        log_Z = -self._log_p_star(x, dt.type(0.))                                    # :708
        beta = delta
        while beta < dt.type(1.) - delta + dt.type(1e-5):                            # :710-711 (storage-dtype beta)
            log_Z = log_Z + self._log_p_star(x, beta)                                # :715
            x = self._ais_transition(x, dt.type(beta + delta), k, seed, it); it += 1  # :717
            log_Z = log_Z - self._log_p_star(x, beta)                                # :719
            beta = dt.type(beta + delta)
        log_Z = log_Z + self._log_p_star(x, dt.type(1.))                             # :728
        # :731-734 -- `tf.cast(tf.log(2.), dtype)`: the float32 value of log 2, also in a float64 model
        log_Z = log_Z + (self.V + self.Hs[0] + self.Hs[1]) * float(np.float32(np.log(2.)))
        return log_Z

    def close(self):
        pass


def dbm_factory(cfg):
    return OracleDBM(cfg)
