"""CPU: the port (oracle/port.py) against the fixtures generated from the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle (the reference ships no
golden vectors of its own)."""
import torch
import pytest

from oracle import port, synth
from helpers import port_spec, port_cfg, product_template_sd


def _params(mc, sc, requires_grad=False):
    sd = synth.synth_state_dict(product_template_sd(mc, sc), 0)
    if mc["spiking_neuron"]["neuron_type"] == "plif":
        synth.spread_plif_w(sd)          # every ParametricLIFNode's w away from 0, as make_golden.py loaded the reference
    return port.params_from_state_dict(sd, requires_grad)


@pytest.mark.parametrize("nt", ["lif", "psn", "plif"])
def test_small_model_eval_matches_reference(golden, nt):
    g = golden(f"small_{nt}_eval.pt")
    mc, sc = synth.small_config(nt)
    P = _params(mc, sc)
    x = synth.synth_voxels(2, 10, 96, 128)
    with torch.no_grad():
        flows = port.ms_flownet_forward(x, P, port_cfg(mc, sc), port_spec(mc), port.BNMode(False))
    assert len(flows) == len(g["flows"]) == 3
    for a, b in zip(flows, g["flows"]):
        assert a.shape == b.shape
        assert torch.equal(a, b), f"max abs diff {(a - b).abs().max().item()}"


@pytest.mark.parametrize("nt", ["lif", "psn", "plif"])
def test_small_model_train_matches_reference(golden, nt):
    g = golden(f"small_{nt}_train.pt")
    mc, sc = synth.small_config(nt)
    P = _params(mc, sc, requires_grad=True)
    B = 2
    x = synth.synth_voxels(B, 10, 96, 128)
    scales = synth.synth_drop_scales(sc["swin_depths"], B)
    flows = port.ms_flownet_forward(x, P, port_cfg(mc, sc), port_spec(mc), port.BNMode(True), scales)
    for a, b in zip(flows, g["flows"]):
        assert torch.allclose(a, b, rtol=0, atol=1e-5), (a - b).abs().max().item()
    gt, mask = synth.synth_labels(B, 96, 128)
    loss = port.flow_loss(flows, gt, mask)
    assert abs(loss.item() - g["loss"]) <= 1e-6 * max(1.0, abs(g["loss"]))
    loss.backward()
    for k, gref in g["grads"].items():
        got = P[k].grad
        assert got is not None, k
        denom = gref.abs().max().clamp_min(1e-12)
        assert ((got - gref).abs().max() / denom).item() < 1e-4, k
    k = "sttmultires_unet.encoders.swin3d.layers.0.swin_blocks.0.mlp.bn1.norm_layer."
    assert torch.allclose(P[k + "running_mean"], g["running_mean_after"], atol=1e-6)
    assert torch.allclose(P[k + "running_var"], g["running_var_after"], atol=1e-6)


def test_en4_shipped_config_eval_matches_reference(golden):
    """MS_SpikingformerFlowNet_en4, window (2,9,9), 288x384 (smallest legal size), lif v_th 0.1."""
    from oracle import reference_loader as rl
    g = golden("en4_lif_eval.pt")
    mc, sc = rl.default_config("lif", input_size=(288, 384))
    P = _params(mc, sc)
    x = synth.synth_voxels(1, 10, 288, 384)
    with torch.no_grad():
        flows = port.ms_flownet_forward(x, P, port_cfg(mc, sc), port_spec(mc), port.BNMode(False))
    for a, s in zip(flows, g["flows"]):
        assert tuple(a.shape) == s["shape"]
        assert torch.equal(a[..., ::8, ::8], s["sub"])
        assert abs(a.double().sum().item() - s["sum"]) <= 1e-6 * max(1.0, abs(s["sum"]))
        assert abs(a.double().abs().sum().item() - s["abs"]) <= 1e-6 * s["abs"]


@pytest.mark.parametrize("name", ["t5_w288", "t10_w466"])
def test_cfg4_shapes_match_reference(golden, name):
    """BASELINE.json configs[3]: 5 time bins with window (2,8,8) at 256x256, and a temporal window of 4."""
    from oracle import reference_loader as rl
    g = golden("cfg4_lif_eval.pt")[name]
    kw = synth.CFG4[name]
    mc, sc = rl.default_config("lif", **kw)
    P = _params(mc, sc)
    x = synth.synth_voxels(1, kw["num_bins"], *kw["input_size"])
    with torch.no_grad():
        flows = port.ms_flownet_forward(x, P, port_cfg(mc, sc), port_spec(mc), port.BNMode(False))
    for a, s in zip(flows, g["flows"]):
        assert tuple(a.shape) == s["shape"]
        assert torch.equal(a[..., ::8, ::8], s["sub"])
        assert abs(a.double().sum().item() - s["sum"]) <= 1e-6 * max(1.0, abs(s["sum"]))


@pytest.mark.parametrize("variant", ["bn", "sdsa"])
@pytest.mark.parametrize("train", [False, True])
def test_sew_stage_matches_reference(golden, variant, train):
    """Spiking_Swin_BasicLayer with Q K^T V attention (Spiking_BN / SDSA), shifted + unshifted block."""
    g = golden("sew_stage.pt")[f"{variant}_{'train' if train else 'eval'}"]
    P, x, cfg, spec = sew_stage_inputs(variant)
    with torch.no_grad():
        xo, xb = port.basic_layer(x.permute(0, 2, 3, 4, 1).contiguous(), P, "L", cfg, 0, spec, port.BNMode(train))
    assert torch.allclose(xb, g["x_pre"], atol=2e-5), (xb - g["x_pre"]).abs().max().item()
    assert torch.allclose(xo.permute(0, 4, 1, 2, 3), g["x_out"], atol=2e-5)


def sew_stage_inputs(variant):
    """State dict / input recipe of oracle/make_golden.py:golden_sew_stage."""
    from sdformerflow_b200.STSwinNet_SNN import Spiking_swin_transformer3D as prod
    kw = {"num_steps": 4, "v_reset": None, "v_th": 0.3, "neuron_type": "lif", "surrogate_fun": "surrogate.ATan()",
          "tau": 2.0, "detach_reset": True, "spike_norm": "BN"}

    class Blk(prod.Spiking_SwinTransformerBlock3D):
        attn_module = prod.Spiking_BN_WindowAttention3D if variant == "bn" else prod.SDSA_WindowAttention3D

    class Lyr(prod.Spiking_Swin_BasicLayer):
        swin_block_type = Blk
    lyr = Lyr(dim=64, input_resolution=(7, 10), depth=2, num_heads=2, window_size=(2, 3, 4),
              pretrained_window_size=(0, 0, 0), mlp_ratio=4.0, version="swinv1", qk_scale=0.125, drop_path=[0.0, 0.0],
              norm_layer="BN", downsample=prod.SpikingPatchMerging, **kw)
    sd = synth.synth_state_dict(lyr.state_dict(), seed=3)
    P = port.params_from_state_dict({"L." + k: v for k, v in sd.items()})
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(2, 64, 4, 7, 10, generator=g) < 0.3).float() * torch.randint(1, 3, (2, 64, 4, 7, 10), generator=g)
    cfg = port.SwinCfg(window_size=(2, 3, 4), depths=(2, 2), num_heads=(2, 4), embed_dim=64, family="sew", attn=variant)
    spec = port.NeuronSpec(4, "lif", 0.3, None, 2.0, True)
    return P, x, cfg, spec
