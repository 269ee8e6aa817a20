"""sklearn-style parameter plumbing
(/root/reference/boltzmann_machines/base/base_model.py:8-67)."""
import copy

import numpy as np

from .base import is_param_name, is_attribute_name
from .mixin import SeedMixin
from ..utils.utils import write_during_training

MAX_SERIALIZED_ELEMS = 1e6


class BaseModel(SeedMixin):
    def __init__(self, *args, **kwargs):
        super(BaseModel, self).__init__(*args, **kwargs)

    def get_params(self, deep=True, include_attributes=True):
        """Public parameters (and, optionally, trailing-underscore attributes)."""
        out = {}
        for k, v in vars(self).items():
            if is_param_name(k) or (include_attributes and is_attribute_name(k)):
                out[k] = copy.deepcopy(v) if deep else v
        return out

    def set_params(self, **params):
        for k, v in params.items():
            known = (is_param_name(k) or is_attribute_name(k)) and hasattr(self, k)
            if not known:
                raise ValueError("invalid param name '{0}'".format(k))
            setattr(self, k, v)
        return self

    def _serialize(self, params):
        """ndarray -> list for JSON; arrays above 1e6 elements are dropped."""
        for k in list(params):
            v = params[k]
            if isinstance(v, np.ndarray):
                if v.size > MAX_SERIALIZED_ELEMS:
                    write_during_training(
                        "WARNING: parameter `{0}` won't be serialized because it is too large:"
                        " ({1:.2f} > 1 Mio elements)".format(k, 1e-6 * v.size))
                    params[k] = None
                else:
                    params[k] = v.tolist()
            elif isinstance(v, (np.floating, np.integer, np.bool_)):
                params[k] = v.item()
            elif isinstance(v, (list, tuple)):
                params[k] = [x.item() if isinstance(x, (np.floating, np.integer, np.bool_)) else x
                             for x in v]
        return params

    def _deserialize(self, params):
        return params
