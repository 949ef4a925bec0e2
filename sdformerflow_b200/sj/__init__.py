"""Product-side mirror of the ``spikingjelly.activation_based`` protocol the reference drives
(functional.reset_net / set_step_mode / set_backend, neuron.* types, layer.* wrappers,
surrogate.* objects) with the neurons backed by the sm_100a kernels.  No cupy, no torch loop.
"""
from . import base, surrogate, functional, layer, neuron  # noqa: F401
