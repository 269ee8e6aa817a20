"""Constructor-chain endpoints shared by every model class.

The model classes use cooperative multiple inheritance: each ``__init__`` takes the keyword arguments it knows
and hands the rest up the MRO.  The three classes here sit at the top of that chain and keep the reference's
observable behaviour (/root/reference/boltzmann_machines/base/mixin.py:7-35): a keyword nobody consumed is a
misspelt parameter (``AttributeError``), ``dtype`` names the storage type, ``random_seed`` seeds the model's
host RNG, from which every public call draws the seed that keys its Philox streams.
"""
import numpy as np

from ..utils import RNG

_DTYPES = {'float32': np.float32, 'float64': np.float64}


def _leftovers(args, kwargs):
    """Positional or keyword arguments that reached the top of the chain were understood by no class."""
    if len(args) or len(kwargs):
        raise AttributeError('Invalid parameters: {0}, {1}'.format(args, kwargs))


class BaseMixin(object):
    def __init__(self, *args, **kwargs):
        _leftovers(args, kwargs)
        super(BaseMixin, self).__init__()


class DtypeMixin(BaseMixin):
    """``dtype`` in {'float32', 'float64'}: storage type of variables, batches and results."""

    def __init__(self, dtype='float32', *args, **kwargs):
        self.dtype = dtype
        super(DtypeMixin, self).__init__(*args, **kwargs)

    _np_dtype = property(lambda self: _DTYPES.get(self.dtype) or getattr(np, self.dtype))
    # the reference's name for the same thing; here it is what the engine is configured with
    _tf_dtype = property(lambda self: self.dtype)


class SeedMixin(BaseMixin):
    """Owns the host-side generator ``_rng`` (numpy MT19937 behind utils.RNG, persisted in random_state.json)."""

    def __init__(self, random_seed=None, *args, **kwargs):
        self.random_seed = random_seed
        self._rng = RNG(seed=random_seed)
        super(SeedMixin, self).__init__(*args, **kwargs)

    def make_random_seed(self):
        """The next per-call seed: one 31-bit draw from the model's generator."""
        draw = self._rng.randint(2 ** 31 - 1)
        return int(draw)
