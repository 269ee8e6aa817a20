"""User-defined stochastic layers (the plugin surface of layers.py:8-36, consumed through v_layer_cls / h_layer_cls): a layer
that is none of the built-in kinds runs on the host-driven engine (boltzmann_machines/_plugin.py) -- GEMMs on the tensor cores
through the C-ABI, the layer's own activation / _sample on the host in between."""
import numpy as np
import pytest

from boltzmann_machines import _native
from boltzmann_machines.layers import BaseLayer, BernoulliLayer, _HostDistribution
from boltzmann_machines.rbm.base_rbm import BaseRBM

pytestmark = pytest.mark.gpu


class MyLogisticLayer(BaseLayer):
    """a user's restatement of logistic units: no `kind`, so nothing of the fused engine knows it"""
    def init(self, batch_size, random_seed=None):
        return np.random.RandomState(random_seed).uniform(size=(batch_size, self.n_units)).astype(self._np_dtype)

    def activation(self, x, b):
        return 1. / (1. + np.exp(-(np.asarray(x) + b)))

    def _sample(self, means):
        means = np.asarray(means)
        return _HostDistribution(lambda g: g.uniform(size=means.shape) < means)


class ClippedLinearLayer(MyLogisticLayer):
    """something the built-in kinds do not offer: means clipped to [0, 1]"""
    def activation(self, x, b):
        return np.clip(0.5 + 0.25 * (np.asarray(x) + b), 0., 1.)


def make(v_cls, h_cls, path, **kw):
    args = dict(n_visible=48, n_hidden=24, v_layer_cls=v_cls, h_layer_cls=h_cls, batch_size=16, max_epoch=2, random_seed=7, verbose=False,
                n_gibbs_steps=2, learning_rate=0.05, momentum=0.5, sample_v_states=False, sample_h_states=False, l2=1e-4,
                metrics_config=dict(msre=True, pll=False, feg=False, train_metrics_every_iter=2), model_path=path)
    args.update(kw)
    return BaseRBM(**args)


def test_custom_layer_model_trains_and_matches_the_builtin_kind_it_restates(tmp_path):
    X = (np.random.RandomState(1).rand(64, 48) < 0.3).astype(np.float32)
    a = make(MyLogisticLayer, MyLogisticLayer, str(tmp_path / 'a') + '/')
    b = make(BernoulliLayer, BernoulliLayer, str(tmp_path / 'b') + '/')
    a.fit(X), b.fit(X)
    assert type(a._engine).__name__ == 'HostLayerRBM' if a._engine is not None else True
    Wa, Wb = a.get_tf_params('weights'), b.get_tf_params('weights')
    for k in ('W', 'vb', 'hb'):
        # deterministic chains (no sampling): the same bf16 GEMMs, fp32 everything else
        np.testing.assert_allclose(Wa[k], Wb[k], atol=2e-3, err_msg=k)
    assert np.abs(Wa['W']).max() > 0.011
    H = a.transform(X)
    assert H.shape == (64, 24) and np.all((H >= 0) & (H <= 1))
    np.testing.assert_allclose(H, b.transform(X), atol=2e-2)


def test_custom_layer_with_its_own_activation_and_sampling(tmp_path):
    X = (np.random.RandomState(2).rand(64, 48) < 0.3).astype(np.float32)
    m = make(BernoulliLayer, ClippedLinearLayer, str(tmp_path / 'c') + '/', sample_h_states=True)
    m.fit(X)
    W = m.get_tf_params('weights')['W']
    assert np.all(np.isfinite(W)) and np.abs(W).max() > 0.011
    m2 = make(BernoulliLayer, ClippedLinearLayer, str(tmp_path / 'd') + '/', sample_h_states=True)
    m2.fit(X)
    np.testing.assert_array_equal(W, m2.get_tf_params('weights')['W'])            # seeded: reproducible
    with pytest.raises(NotImplementedError, match='free energy'):
        make(BernoulliLayer, ClippedLinearLayer, str(tmp_path / 'e') + '/', metrics_config=dict(msre=True, pll=True, train_metrics_every_iter=1)).fit(X)
