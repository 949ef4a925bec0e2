"""Shared builders for the parity tests (oracle side = oracle.port, product side = sdformerflow_b200)."""
import copy

import torch

from oracle import port, synth


def port_spec(mc):
    sn = mc["spiking_neuron"]
    return port.NeuronSpec(sn["num_steps"], sn["neuron_type"], sn["v_th"], sn["v_reset"], sn["tau"], sn["detach_reset"])


def port_cfg(mc, sc, attn="qk", family="ms"):
    swin = port.SwinCfg(window_size=sc["window_size"], depths=sc["swin_depths"], num_heads=sc["swin_num_heads"],
                        embed_dim=sc["base_num_channels"], family=family, attn=attn)
    return port.FlowNetCfg(swin, num_bins=mc["num_bins"], num_steps=mc["spiking_neuron"]["num_steps"])


def build_product(mc, sc, device="cpu", train=False, seed=0):
    """Product model with the synthetic weights of oracle.synth (same recipe as the goldens)."""
    from sdformerflow_b200.STSwinNet_SNN import Spiking_STSwinNet as prod
    from sdformerflow_b200.sj import functional
    model = getattr(prod, mc["name"])(copy.deepcopy(mc), copy.deepcopy(sc))
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed), strict=True)
    functional.set_step_mode(model, "m")
    model.train(train)
    return model.to(device)


def product_template_sd(mc, sc):
    from sdformerflow_b200.STSwinNet_SNN import Spiking_STSwinNet as prod
    return getattr(prod, mc["name"])(copy.deepcopy(mc), copy.deepcopy(sc)).state_dict()


def flip_rate(a, b):
    return (a != b).float().mean().item()


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def epe(a, b):
    """mean end-point error in px between two (B,2,H,W) flows"""
    return (a - b).pow(2).sum(1).sqrt().mean().item()
