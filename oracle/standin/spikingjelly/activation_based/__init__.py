from . import base, surrogate, functional, layer, neuron  # noqa: F401
