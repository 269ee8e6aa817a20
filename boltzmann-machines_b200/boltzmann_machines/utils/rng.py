"""Host-side random state: a ``numpy.random.RandomState`` whose state survives
a JSON round trip.  Mirrors /root/reference/boltzmann_machines/utils/rng.py:4-62
(same class name, same ``reseed/get_state/set_state`` contract) so that
``random_state.json`` files are interchangeable.
"""
import numpy as np


class RNG(np.random.RandomState):
    """Seeded generator.  ``RNG(None)`` is OS-seeded; ``RNG(int)`` is reproducible.

    >>> g = RNG(1337)
    >>> snap = g.get_state()
    >>> float(g.rand())
    0.2620246750155817
    >>> float(g.rand())
    0.1586839721544656
    >>> float(g.reseed().rand())
    0.2620246750155817
    >>> import json
    >>> float(g.set_state(json.loads(json.dumps(snap))).rand())
    0.2620246750155817
    """

    def __init__(self, seed=None):
        np.random.RandomState.__init__(self, seed)
        self._seed = seed

    def reseed(self):
        """Rewind to the construction seed (no-op for an unseeded generator)."""
        if self._seed is not None:
            self.seed(self._seed)
        return self

    def get_state(self):
        """MT19937 state with the key vector as a plain list (JSON friendly)."""
        name, key, pos, has_gauss, cached = np.random.RandomState.get_state(self)
        return [name, [int(w) for w in key], int(pos), int(has_gauss), float(cached)]

    def set_state(self, state):
        name, key, pos, has_gauss, cached = state
        np.random.RandomState.set_state(
            self, (str(name), np.asarray(key, dtype=np.uint32), int(pos), int(has_gauss), float(cached)))
        return self
