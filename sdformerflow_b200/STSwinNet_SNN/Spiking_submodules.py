"""Extra neurons of reference models/STSwinNet_SNN/Spiking_submodules.py.

PSN (:183-215) is built on the K1p kernel.  GatedLIFNode (:94-180) and SLTTLIFNode (:11-91) are
not used by any shipped config and are out of scope for the hot path: the names exist so that
``from ...Spiking_submodules import *`` keeps working, constructing them raises.
"""
import math
import torch
from torch import nn
from ..sj import base, surrogate
from .. import capi, ops

__all__ = ["PSN", "GatedLIFNode", "SLTTLIFNode"]


class PSN(nn.Module, base.MultiStepModule):
    """Parallel spiking neuron: s = heaviside(W x + b), W in R^{TxT}, b in R^{Tx1} (init -1)."""

    def __init__(self, T: int, surrogate_function=None):
        super().__init__()
        self.T = T
        self.surrogate_function = surrogate_function if surrogate_function is not None else surrogate.ATan()
        self.weight = nn.Parameter(torch.zeros([T, T]))
        self.bias = nn.Parameter(torch.zeros([T, 1]))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        nn.init.constant_(self.bias, -1.0)

    def neuron_cfg(self):
        sf = self.surrogate_function
        return ops.NeuronCfg(kind=capi.SDF_NEURON_IF, v_th=0.0, v_reset=None,
                             surrogate=getattr(sf, "kind", capi.SDF_SG_ATAN), sg_alpha=float(getattr(sf, "alpha", 2.0)))

    def forward(self, x_seq):
        return ops.psn(x_seq, self.weight, self.bias, self.neuron_cfg(), 0)

    def extra_repr(self):
        return f"T={self.T}, "


class GatedLIFNode(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("GatedLIFNode (glif) is outside the B200 hot-path scope (no shipped config uses it)")


class SLTTLIFNode(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("SLTTLIFNode is outside the B200 hot-path scope (no shipped config uses it)")
