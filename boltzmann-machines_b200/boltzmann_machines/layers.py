"""Stochastic-unit plugins: the objects a model is parameterised with through
``v_layer_cls / v_layer_params / h_layer_cls / h_layer_params``.

Same surface as /root/reference/boltzmann_machines/layers.py:8-89
(``BaseLayer(n_units, dtype)`` with ``init``, ``activation``, ``_sample``,
``sample``; ``BernoulliLayer``, ``MultinomialLayer(n_samples)``,
``GaussianLayer(sigma)``).  In this engine a layer is a *description*: the
fused CUDA epilogue implements the three built-in unit kinds, selected by
``layer.kind`` and parameterised by ``layer.engine_params()``.  The numpy
methods below state each kind's semantics on host arrays (they document the
plugin contract and serve small host-side needs such as drawing initial
particles); they are not on the training path.
"""
import numpy as np

from .base import DtypeMixin


class _HostDistribution(object):
    """Tiny stand-in for the object ``_sample`` returns in the reference
    (something with a ``.sample()``)."""
    def __init__(self, draw):
        self._draw = draw

    def sample(self, rng=None):
        return self._draw(rng or np.random)


class BaseLayer(DtypeMixin):
    """One layer of stochastic units."""
    kind = None            # 'bernoulli' | 'multinomial' | 'gaussian'

    def __init__(self, n_units, *args, **kwargs):
        super(BaseLayer, self).__init__(*args, **kwargs)
        self.n_units = n_units

    def engine_params(self):
        """Scalars/arrays the CUDA epilogue needs besides ``kind``."""
        return {}

    def init(self, batch_size, random_seed=None):
        """Random initial states, shape (batch_size, n_units)."""
        raise NotImplementedError('`init` is not implemented')

    def activation(self, x, b):
        """Means given total input ``x`` (bias excluded) and bias ``b``."""
        raise NotImplementedError('`activation` is not implemented')

    def _sample(self, means):
        raise NotImplementedError('`sample` is not implemented')

    def sample(self, means, rng=None):
        return np.asarray(self._sample(means).sample(rng), dtype=self._np_dtype)


class BernoulliLayer(BaseLayer):
    kind = 'bernoulli'

    def init(self, batch_size, random_seed=None):
        return np.random.RandomState(random_seed).uniform(
            size=(batch_size, self.n_units)).astype(self._np_dtype)

    def activation(self, x, b):
        return 1. / (1. + np.exp(-(np.asarray(x) + b)))

    def _sample(self, means):
        means = np.asarray(means)
        return _HostDistribution(lambda g: g.uniform(size=means.shape) < means)


class MultinomialLayer(BaseLayer):
    kind = 'multinomial'

    def __init__(self, n_samples=100, *args, **kwargs):
        super(MultinomialLayer, self).__init__(*args, **kwargs)
        self.n_samples = float(n_samples)

    def engine_params(self):
        return {'n_samples': self.n_samples}

    def init(self, batch_size, random_seed=None):
        t = np.random.RandomState(random_seed).uniform(size=(batch_size, self.n_units))
        return (t / t.sum()).astype(self._np_dtype)

    def activation(self, x, b):
        z = np.asarray(x) + b
        e = np.exp(z - z.max(axis=-1, keepdims=True))
        return self.n_samples * e / e.sum(axis=-1, keepdims=True)

    def _sample(self, means):
        means = np.atleast_2d(np.asarray(means, dtype=np.float64))
        p = means / means.sum(axis=1, keepdims=True)
        n = int(self.n_samples)
        return _HostDistribution(lambda g: np.stack([g.multinomial(n, row) for row in p]))


class GaussianLayer(BaseLayer):
    kind = 'gaussian'

    def __init__(self, sigma, *args, **kwargs):
        super(GaussianLayer, self).__init__(*args, **kwargs)
        self.sigma = np.asarray(sigma)

    def engine_params(self):
        return {'sigma': self.sigma}

    def init(self, batch_size, random_seed=None):
        t = np.random.RandomState(random_seed).normal(size=(batch_size, self.n_units))
        return (t * self.sigma).astype(self._np_dtype)

    def activation(self, x, b):
        return np.asarray(x) * self.sigma + b

    def _sample(self, means):
        means = np.asarray(means)
        return _HostDistribution(lambda g: means + self.sigma * g.normal(size=means.shape))
