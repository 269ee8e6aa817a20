"""Cooperative mixins: kwargs hygiene, dtype, seeding
(/root/reference/boltzmann_machines/base/mixin.py:7-35)."""
import numpy as np

from ..utils import RNG


class BaseMixin(object):
    """End of every cooperative ``__init__`` chain: anything still unconsumed
    is a misspelt parameter."""
    def __init__(self, *args, **kwargs):
        if args or kwargs:
            raise AttributeError('Invalid parameters: {0}, {1}'.format(args, kwargs))
        super(BaseMixin, self).__init__()


class DtypeMixin(BaseMixin):
    def __init__(self, dtype='float32', *args, **kwargs):
        super(DtypeMixin, self).__init__(*args, **kwargs)
        self.dtype = dtype

    @property
    def _np_dtype(self):
        return getattr(np, self.dtype)

    # the reference exposes ``_tf_dtype``; here it names the engine's storage type
    @property
    def _tf_dtype(self):
        return self.dtype


class SeedMixin(BaseMixin):
    def __init__(self, random_seed=None, *args, **kwargs):
        super(SeedMixin, self).__init__(*args, **kwargs)
        self.random_seed = random_seed
        self._rng = RNG(seed=self.random_seed)

    def make_random_seed(self):
        """Next per-call seed (keys the Philox streams of one public call)."""
        return int(self._rng.randint(2 ** 31 - 1))
