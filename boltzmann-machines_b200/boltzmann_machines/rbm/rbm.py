"""Concrete RBMs: which unit kinds sit on each side, plus model-specific bits
(/root/reference/boltzmann_machines/rbm/rbm.py:10-123).  Free energies
(Bernoulli :17-22, Multinomial :50-60, Gaussian :109-116) are evaluated inside
the engine; the choice follows from the layer kinds."""
import numpy as np

from .base_rbm import BaseRBM
from ..layers import BernoulliLayer, MultinomialLayer, GaussianLayer


class BernoulliRBM(BaseRBM):
    """Bernoulli visible and hidden units."""
    def __init__(self, model_path='b_rbm_model/', *args, **kwargs):
        super(BernoulliRBM, self).__init__(v_layer_cls=BernoulliLayer,
                                           h_layer_cls=BernoulliLayer,
                                           model_path=model_path, *args, **kwargs)


class MultinomialRBM(BaseRBM):
    """Bernoulli visible units and one Multinomial hidden unit, i.e.
    ``n_samples`` softmax units over ``n_hidden`` states with tied weights."""
    def __init__(self, n_samples=100, model_path='m_rbm_model/', *args, **kwargs):
        self.n_samples = n_samples
        super(MultinomialRBM, self).__init__(v_layer_cls=BernoulliLayer,
                                             h_layer_cls=MultinomialLayer,
                                             h_layer_params=dict(n_samples=self.n_samples),
                                             model_path=model_path, *args, **kwargs)

    def transform(self, *args, **kwargs):
        H = super(MultinomialRBM, self).transform(*args, **kwargs)
        H /= float(self.n_samples)          # expected counts -> probabilities
        return H


class GaussianRBM(BaseRBM):
    """Gaussian visible (fixed ``sigma``) and Bernoulli hidden units.  Inputs
    should be zero-mean; with unit-variance inputs use ``sigma=1``."""
    _SCOPES = dict(BaseRBM._SCOPES, input_data=('sigma',))

    def __init__(self, learning_rate=1e-3, sigma=1., model_path='g_rbm_model/', *args, **kwargs):
        self.sigma = sigma
        super(GaussianRBM, self).__init__(v_layer_cls=GaussianLayer,
                                          v_layer_params=dict(sigma=self.sigma),
                                          h_layer_cls=BernoulliLayer,
                                          learning_rate=learning_rate,
                                          model_path=model_path, *args, **kwargs)
        if hasattr(self.sigma, '__iter__'):
            self.sigma = np.asarray(self.sigma)


def logit_mean(X):
    """log(p / (1 - p)) of the per-feature mean, clipped away from 0 and 1:
    the recommended visible-bias initialisation."""
    p = np.clip(np.mean(X, axis=0), 1e-7, 1. - 1e-7)
    return np.log(p) - np.log1p(-p)
