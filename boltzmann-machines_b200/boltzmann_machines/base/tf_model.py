"""Import-compatibility alias: code written against the reference imports
``boltzmann_machines.base.tf_model.TensorFlowModel`` / ``run_in_tf_session``
(/root/reference/boltzmann_machines/base/tf_model.py).  Both now resolve to the
native-engine shim; no TensorFlow is involved."""
from .native_model import (NativeModel, TensorFlowModel, run_in_tf_session,   # noqa: F401
                           set_engine_factory, get_engine_factory)
