"""Parameter plumbing of the model classes, sklearn style: what counts as a parameter is decided by its NAME
(`base.is_param_name`: public, no trailing underscore; `base.is_attribute_name`: fitted state, trailing
underscore), and `params.json` is the JSON image of those names.  Behaviour follows
/root/reference/boltzmann_machines/base/base_model.py:8-67 (same names, same errors, same size limit)."""
import copy

import numpy as np

from .base import is_param_name, is_attribute_name
from .mixin import SeedMixin
from ..utils.utils import write_during_training

MAX_SERIALIZED_ELEMS = 1e6
_NP_SCALARS = (np.floating, np.integer, np.bool_)


def _plain(value):
    """numpy scalar -> python scalar (json.dump rejects the former)."""
    return value.item() if isinstance(value, _NP_SCALARS) else value


class BaseModel(SeedMixin):
    def __init__(self, *args, **kwargs):
        super(BaseModel, self).__init__(*args, **kwargs)

    def _is_visible_name(self, name, include_attributes=True):
        return is_param_name(name) or (include_attributes and is_attribute_name(name))

    def get_params(self, deep=True, include_attributes=True):
        """{name: value} of the constructor parameters and, unless told otherwise, of the fitted attributes."""
        out = {}
        for name, value in vars(self).items():
            if self._is_visible_name(name, include_attributes):
                out[name] = copy.deepcopy(value) if deep else value
        return out

    def set_params(self, **params):
        """Assign existing parameters / attributes in the order given; an unknown name is an error (names before it
        have been assigned by then, as in the reference) and nothing is ever created."""
        for k in params:
            if not (self._is_visible_name(k) and hasattr(self, k)):
                raise ValueError("invalid param name '{0}'".format(k))
            setattr(self, k, params[k])
        return self

    def _serialize(self, params):
        """In place: arrays become lists (dropped with a warning above 1e6 elements), numpy scalars become python ones."""
        for k in list(params):
            v = params[k]
            if isinstance(v, np.ndarray):
                too_large = v.size > MAX_SERIALIZED_ELEMS
                if too_large:
                    write_during_training("WARNING: parameter `{0}` won't be serialized because it is too large:"
                                          " ({1:.2f} > 1 Mio elements)".format(k, 1e-6 * v.size))
                params[k] = None if too_large else v.tolist()
            elif isinstance(v, (list, tuple)):
                params[k] = [_plain(x) for x in v]
            else:
                params[k] = _plain(v)
        return params

    def _deserialize(self, params):
        return params
