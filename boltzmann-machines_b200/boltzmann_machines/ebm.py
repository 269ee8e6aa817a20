"""Root of the model hierarchy: anything that assigns an energy to joint states of visible and latent units
(/root/reference/boltzmann_machines/ebm.py:4-17 plays this role on top of its TensorFlow base class)."""
from .base import NativeModel


class EnergyBasedModel(NativeModel):
    def _free_energy(self, v):
        """Mean free energy of the rows of ``v``.  Concrete models do not implement it on the host: the engine
        evaluates it (`bm_rbm_metrics` with BM_METRIC_FREE_ENERGY)."""
        raise NotImplementedError('`free_energy` is not implemented')
