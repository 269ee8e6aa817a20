"""Engine for models parameterised with USER-DEFINED stochastic layers (layers.py:8-36: a `BaseLayer` subclass that is not
one of the built-in kinds; consumed through `v_layer_cls / h_layer_cls`, rbm/base_rbm.py:96-117).

The fused CUDA epilogues implement Bernoulli, Gaussian and multinomial units.  A custom layer states its `activation(x, b)` and
`_sample(means)` on host arrays (the contract of `layers.BaseLayer` here), so its model runs on the *slow host-driven* path the
plugin surface promises: every GEMM of the CD-k step -- v W, h W^T and X^T h0 - v_k^T h_k -- runs on the
GPU's tensor cores through the C-ABI (`bm_debug_tc_gemm`: bf16 operands, fp32 accumulation), the user's two callbacks and the
momentum update run on the host between them.  It is a plugin path, not a CPU fallback: it needs the GPU like every other
engine, no built-in unit kind ever takes it, and it is orders of magnitude slower than the fused path (host round trip per
half-step).  Metrics that need a free energy (`pll`, `free_energy`, `feg`) are not defined for an arbitrary layer.
"""
import numpy as np

from . import _native


class HostLayerRBM(object):
    compute = 'bf16-plugin'

    def __init__(self, cfg, v_layer, h_layer, ctx=None):
        self.cfg = dict(cfg)
        self.ctx = ctx or _native.Context.default()
        self.V, self.H = int(cfg['n_visible']), int(cfg['n_hidden'])
        self.dt = np.dtype(cfg.get('dtype', 'float32'))
        self.vl, self.hl = v_layer, h_layer
        z = lambda *s: np.zeros(s, dtype=self.dt)
        self.p = dict(W=z(self.V, self.H), vb=z(self.V), hb=z(self.H), dW=z(self.V, self.H), dvb=z(self.V), dhb=z(self.H),
                      q_means=z(self.H))
        self.m_up = 2.0 if cfg.get('dbm_first', False) else 1.0          # base_rbm.py:256-262
        self.m_dn = 2.0 if cfg.get('dbm_last', False) else 1.0

    # -- variables -------------------------------------------------------------------------
    def set_params(self, d):
        for k, v in d.items():
            if k in self.p:
                self.p[k] = np.array(v, dtype=self.dt).reshape(self.p[k].shape)

    def get_params(self, names=None):
        return {k: self.p[k].copy() for k in (names or list(self.p)) if k in self.p}

    def init_normal_W(self, stddev, op_seed):
        eng = _native.CudaRBM(dict(n_visible=self.V, n_hidden=self.H, dtype=self.dt.name, compute='fp32'), ctx=self.ctx)
        eng.init_normal_W(stddev, op_seed)                                # the reference's seeded stream (test_rbm.py:65-67)
        self.p['W'] = eng.get_params(['W'])['W']
        eng.close()

    # -- the three GEMMs, on the tensor cores ------------------------------------------------
    def _up(self, v):
        x = _native.debug_tc_gemm(v, self.p['W'], a_t=False, b_t=True, ctx=self.ctx)           # v W: W stored [K, N]
        return np.asarray(self.hl.activation(self.m_up * x, self.m_up * self.p['hb']), dtype=self.dt)

    def _down(self, h):
        x = _native.debug_tc_gemm(h, self.p['W'], a_t=False, b_t=False, ctx=self.ctx)          # h W^T: W stored [N, K]
        return np.asarray(self.vl.activation(self.m_dn * x, self.m_dn * self.p['vb']), dtype=self.dt)

    def _draw(self, layer, means, seed, site, t, tick):
        rng = np.random.RandomState([int(seed) & 0xffffffff, (int(seed) >> 32) & 0xffffffff, site, t, int(tick) & 0xffffffff])
        return np.asarray(layer._sample(means).sample(rng), dtype=self.dt)

    def _prepare(self, X, seed, tick):
        X = np.ascontiguousarray(X, dtype=self.dt)
        keep = self.cfg.get('dropout', None)
        if keep is not None:                                              # base_rbm.py:417-418
            rng = np.random.RandomState([int(seed) & 0xffffffff, 0, 0, 0, int(tick) & 0xffffffff])
            X = X / self.dt.type(keep) * np.floor(keep + rng.uniform(size=X.shape)).astype(self.dt)
        return X

    def _chain(self, X, k, seed, tick):
        sh, sv = self.cfg.get('sample_h', True), self.cfg.get('sample_v', False)
        h0 = self._up(X)                                                  # base_rbm.py:421-426
        h = self._draw(self.hl, h0, seed, 1, 0, tick) if sh else h0
        v = vm = hm = None
        for t in range(1, int(k) + 1):                                    # base_rbm.py:367-384
            vm = self._down(h)
            v = self._draw(self.vl, vm, seed, 2, t, tick) if sv else vm
            hm = self._up(v)
            h = self._draw(self.hl, hm, seed, 3, t, tick) if (sh and t < int(k)) else hm
        return h0, v, vm, hm

    def _metrics(self, names, X, vm):
        out = {}
        for n in names:
            if n == 'msre':
                out[n] = float(np.mean((X - vm) ** 2))                    # base_rbm.py:486-488
            elif n == 'l2_loss':
                out[n] = float(self.cfg.get('l2', 0.) * 0.5 * np.sum(self.p['W'].astype(np.float64) ** 2))
            else:
                raise NotImplementedError("metric '{0}' needs a free energy, which a user-defined layer does not define".format(n))
        return out

    # -- engine interface --------------------------------------------------------------------
    def train_step(self, X, lr, momentum, k, seed, tick, metrics=()):
        X = self._prepare(X, seed, tick)
        h0, v, vm, hm = self._chain(X, k, seed, tick)
        res = self._metrics(metrics, X, vm) if metrics else None
        N = self.dt.type(X.shape[0])
        p, c = self.p, self.cfg
        G = _native.debug_tc_gemm(X, h0, a_t=True, b_t=True, A2=v, B2=hm, neg2=True, ctx=self.ctx)     # base_rbm.py:447-448
        dW = G.astype(self.dt) / N - self.dt.type(c.get('l2', 0.)) * p['W']
        dvb = np.mean(X - v, axis=0)
        dhb = np.mean(h0 - hm, axis=0)
        q = self.dt.type(c.get('sparsity_damping', 0.9)) * p['q_means'] + \
            self.dt.type(1 - c.get('sparsity_damping', 0.9)) * hm.sum(axis=0)                           # :457-459
        pen = self.dt.type(c.get('sparsity_cost', 0.)) * (q - self.dt.type(c.get('sparsity_target', 0.1)))
        p['q_means'] = q.astype(self.dt)
        dhb = dhb - pen
        dW = dW - pen[None, :]
        lr, mom = self.dt.type(lr), self.dt.type(momentum)
        p['dW'] = (lr * (mom * p['dW'] + dW)).astype(self.dt); p['W'] = p['W'] + p['dW']                # :465-474
        p['dvb'] = (lr * (mom * p['dvb'] + dvb)).astype(self.dt); p['vb'] = p['vb'] + p['dvb']
        p['dhb'] = (lr * (mom * p['dhb'] + dhb)).astype(self.dt); p['hb'] = p['hb'] + p['dhb']
        return res

    def transform(self, X, k, seed, tick):
        X = self._prepare(X, seed, tick)
        return self._chain(X, k, seed, tick)[3]

    def metrics(self, X, k, seed, tick, names):
        X = self._prepare(X, seed, tick)
        vm = self._chain(X, k, seed, tick)[2] if 'msre' in names else None
        return self._metrics(names, X, vm)

    def close(self):
        pass
