from .base import NativeModel


class EnergyBasedModel(NativeModel):
    """Energy-based model with latent variables
    (/root/reference/boltzmann_machines/ebm.py:4-17)."""
    def __init__(self, *args, **kwargs):
        super(EnergyBasedModel, self).__init__(*args, **kwargs)

    def _free_energy(self, v):
        """Average free energy of the rows of ``v`` (evaluated by the engine)."""
        raise NotImplementedError('`free_energy` is not implemented')
