"""GPU parity tests, block / model level: product modules on cuda:0 vs the CPU oracle port and the
golden fixtures generated from the unmodified reference."""
import pytest
import torch

from oracle import port, synth
from helpers import port_spec, port_cfg, build_product, epe, flip_rate

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reset(model):
    from sdformerflow_b200.sj import functional
    functional.reset_net(model)


@pytest.mark.parametrize("train", [False, True])
@pytest.mark.parametrize("shift", [(0, 0, 0), (1, 1, 2)])
def test_ms_block_fwd_bwd(train, shift):
    """MS_Spiking_SwinTransformerBlock3D (QK-gate attention + MS MLP), pad > 0, shifted and unshifted."""
    from sdformerflow_b200.STSwinNet_SNN import Spiking_swin_transformer3D as prod
    kw = {"num_steps": 10, "v_reset": None, "v_th": 0.1, "neuron_type": "lif", "surrogate_fun": "surrogate.ATan()",
          "tau": 2.0, "detach_reset": True, "spike_norm": "BN"}
    C, nH, B, D, H, W = 96, 3, 2, 10, 7, 10
    blk = prod.MS_Spiking_SwinTransformerBlock3D(C, (H, W), nH, window_size=(2, 3, 4), shift_size=shift, drop_path=0.1,
                                                  norm_layer="BN", qk_scale=0.125, **kw)
    sd = synth.synth_state_dict(blk.state_dict(), seed=2)
    blk.load_state_dict(sd)
    blk.train(train).to(DEV)
    P = port.params_from_state_dict({"b." + k: v for k, v in sd.items()}, requires_grad=True)
    g0 = torch.Generator().manual_seed(3)
    x = (torch.randn(B, D, H, W, C, generator=g0) * 0.5).requires_grad_(True)
    ds = torch.tensor([1.0 / 0.9, 0.0]) if train else None
    cfg = port.SwinCfg(window_size=(2, 3, 4), depths=(2,), num_heads=(nH,), embed_dim=C)
    ref = port.swin_block(x, P, "b", cfg, nH, shift, None, port.NeuronSpec(10, "lif", 0.1, None, 2.0, True),
                          port.BNMode(train), ds)
    go = torch.randn(ref.shape, generator=g0)
    ref.backward(go)
    if train:
        blk.drop_path.forced = ds
    xg = x.detach().to(DEV).requires_grad_(True)
    out = blk(xg)
    out.backward(go.to(DEV))
    err = (out.detach().cpu() - ref.detach()).abs()
    # membrane-potential stream: <= 1e-4 relative except where an upstream spike flipped at a tie
    assert (err > 1e-4 * ref.abs().max()).float().mean().item() <= 2e-3, err.max().item()
    gerr = (xg.grad.cpu() - x.grad).abs()
    assert (gerr > 1e-3 * x.grad.abs().max()).float().mean().item() <= 1e-2
    for name in ("attn.linear_k.weight", "mlp.fc1.weight", "mlp.bn2.norm_layer.bias", "attn.positional_encoding"):
        a, b = dict(blk.named_parameters())[name].grad.cpu(), P["b." + name].grad
        assert ((a - b).abs() > 2e-2 * b.abs().max()).float().mean().item() <= 2e-2, name


@pytest.mark.parametrize("nt", ["lif", "psn"])
def test_small_model_eval_epe(golden, nt):
    """MS 3-encoder model: flow within 1e-3 px EPE of the reference (golden) and of the port."""
    g = golden(f"small_{nt}_eval.pt")
    mc, sc = synth.small_config(nt)
    model = build_product(mc, sc, DEV, train=False)
    x = synth.synth_voxels(2, 10, 96, 128)
    _reset(model)
    with torch.no_grad():
        flows = model(x.to(DEV))["flow"]
    assert len(flows) == 3
    for a, b in zip(flows, g["flows"]):
        assert a.shape == b.shape
        assert epe(a.cpu(), b) <= 1e-3, epe(a.cpu(), b)


def test_small_model_train_step(golden):
    """fwd + loss + bwd in train mode (BN batch stats, injected DropPath masks) vs the reference."""
    g = golden("small_lif_train.pt")
    mc, sc = synth.small_config("lif")
    model = build_product(mc, sc, DEV, train=True)
    B = 2
    x = synth.synth_voxels(B, 10, 96, 128)
    scales = synth.synth_drop_scales(sc["swin_depths"], B)
    blocks = [b for lyr in model.sttmultires_unet.encoders.swin3d.layers for b in lyr.swin_blocks]
    for b, s in zip(blocks, scales):
        if s is not None:
            b.drop_path.forced = s
    _reset(model)
    flows = model(x.to(DEV))["flow"]
    for a, b in zip(flows, g["flows"]):
        assert epe(a.detach().cpu(), b) <= 1e-3
    gt, mask = synth.synth_labels(B, 96, 128)
    loss = port.flow_loss(flows, gt.to(DEV), mask.to(DEV))
    assert abs(loss.item() - g["loss"]) <= 1e-3 * abs(g["loss"])
    loss.backward()
    named = dict(model.named_parameters())
    for k, gref in g["grads"].items():
        got = named[k].grad.cpu()
        rel = (got - gref).norm() / gref.norm().clamp_min(1e-12)
        assert rel.item() <= 5e-2, (k, rel.item())


def test_en4_shipped_config_eval(golden):
    """The shipped model (MS en4, window (2,9,9)) at 288x384 against the reference fixture."""
    from oracle import reference_loader as rl
    g = golden("en4_lif_eval.pt")
    mc, sc = rl.default_config("lif", input_size=(288, 384))
    model = build_product(mc, sc, DEV, train=False)
    x = synth.synth_voxels(1, 10, 288, 384)
    _reset(model)
    with torch.no_grad():
        flows = model(x.to(DEV))["flow"]
    for a, s in zip(flows, g["flows"]):
        sub = a[..., ::8, ::8].cpu()
        assert epe(sub, s["sub"]) <= 1e-3, epe(sub, s["sub"])


def test_double_forward_without_reset_raises():
    mc, sc = synth.small_config("lif")
    model = build_product(mc, sc, DEV, train=False)
    x = synth.synth_voxels(1, 10, 96, 128).to(DEV)
    _reset(model)
    with torch.no_grad():
        model(x)
        with pytest.raises(RuntimeError):
            model(x)
