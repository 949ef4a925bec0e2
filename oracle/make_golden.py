"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, imported through oracle/standin) on the deterministic recipe of oracle/synth.py.

Run in the build container:  python oracle/make_golden.py
The fixtures pin the CPU port (oracle/port.py) — and through it the CUDA path — to the reference's
own outputs; /root/reference is not needed to consume them.
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_loader as rl  # noqa: E402
from oracle import synth  # noqa: E402

rl._activate()  # puts oracle/standin and /root/reference on sys.path

OUT = os.path.join(ROOT, "tests", "golden")


def _load_synth(model, seed=0, spread_plif=False):
    sd = synth.synth_state_dict(model.state_dict(), seed)
    if spread_plif:
        synth.spread_plif_w(sd)
    model.load_state_dict(sd, strict=True)
    return sd


def _summ(t, step=8):
    """Sub-sampled copy + moments: keeps fixtures small while still pinning every region."""
    return {"sub": t[..., ::step, ::step].clone(), "sum": t.double().sum().item(), "abs": t.double().abs().sum().item(),
            "shape": tuple(t.shape)}


def golden_small_model(neuron_type, train):
    """MS 3-encoder model, 96x128, window (2,3,4): flows (+ loss and a few gradients in train mode)."""
    from spikingjelly.activation_based import functional
    mc, sc = synth.small_config(neuron_type)
    model = rl.build_reference_model(mc, sc, seed=0, train=train)
    _load_synth(model, spread_plif=neuron_type == "plif")       # plif: every site's w away from 0 (synth.spread_plif_w)
    B = 2
    x = synth.synth_voxels(B, 10, 96, 128)
    out = {"neuron_type": neuron_type, "train": train}
    functional.reset_net(model)
    if not train:
        with torch.no_grad():
            flows = model(x)["flow"]
        out["flows"] = [f.clone() for f in flows]
        return out
    # train mode: BN batch statistics, DropPath with injected masks (so CPU/GPU RNG need not agree)
    scales = synth.synth_drop_scales(sc["swin_depths"], B)
    blocks = [b for lyr in model.sttmultires_unet.encoders.swin3d.layers for b in lyr.swin_blocks]

    class Forced(torch.nn.Module):
        def __init__(self, s):
            super().__init__()
            self.s = s

        def forward(self, t):
            return t if self.s is None else t * self.s.view(-1, 1, 1, 1, 1)

    for b, s in zip(blocks, scales):
        b.drop_path = Forced(s)
    flows = model(x)["flow"]
    gt, mask = synth.synth_labels(B, 96, 128)
    sys.path.insert(0, rl.REFERENCE_ROOT)
    from loss.flow_supervised import flow_loss_supervised
    crit = flow_loss_supervised({"metrics": {"flow_scaling": 1}, "loss": {"lambda_mod": 1, "lambda_ang": 0}}, "cpu")
    loss = crit(flows, gt, mask)
    loss.backward()
    out["flows"] = [f.detach().clone() for f in flows]
    out["loss"] = loss.item()
    grads = {}
    named = dict(model.named_parameters())
    for k in ["sttmultires_unet.encoders.swin3d.layers.0.swin_blocks.0.attn.linear_q.weight",
              "sttmultires_unet.encoders.swin3d.layers.0.swin_blocks.1.attn.positional_encoding",
              "sttmultires_unet.encoders.swin3d.layers.0.swin_blocks.1.attn.bn_k.norm_layer.weight",
              "sttmultires_unet.encoders.swin3d.layers.1.swin_blocks.0.mlp.fc1.weight",
              "sttmultires_unet.encoders.swin3d.layers.1.swin_blocks.1.attn.proj.bias",
              "sttmultires_unet.encoders.swin3d.layers.0.downsample.reduction.weight",
              "sttmultires_unet.encoders.swin3d.patch_embed.head.conv.0.weight",
              "sttmultires_unet.preds.2.conv.0.weight"] + (
            # plif: the gradient of 1/tau's parameter at a plain, a BN-fused, a window-fused and a QK-gate site
            ["sttmultires_unet.encoders.swin3d.patch_embed.head.sn.spiking_neuron.w",
             "sttmultires_unet.encoders.swin3d.layers.0.swin_blocks.0.mlp.sn1.spiking_neuron.w",
             "sttmultires_unet.encoders.swin3d.layers.0.swin_blocks.0.attn.proj_sn.spiking_neuron.w",
             "sttmultires_unet.encoders.swin3d.layers.0.swin_blocks.1.attn.sn_k.spiking_neuron.w",
             "sttmultires_unet.decoders.1.sn.spiking_neuron.w"] if neuron_type == "plif" else []):
        grads[k] = named[k].grad.clone()
    out["grads"] = grads
    out["running_mean_after"] = model.state_dict()[
        "sttmultires_unet.encoders.swin3d.layers.0.swin_blocks.0.mlp.bn1.norm_layer.running_mean"].clone()
    out["running_var_after"] = model.state_dict()[
        "sttmultires_unet.encoders.swin3d.layers.0.swin_blocks.0.mlp.bn1.norm_layer.running_var"].clone()
    return out


def golden_small_train_sensitivity(n_samples=4, log2_jitter=-21):
    """How far the reference's OWN train-mode loss (same model / input / DropPath masks as small_lif_train.pt) moves when
    every weight matrix is jittered by a relative +-2^-21 (4 ulp): the yardstick for the free-running train-step test.
    The product's arithmetic is a few ulp away from torch's per pre-activation (BatchNorm as one fused affine
    scale*x + shift, spike-GEMM weights in 23-bit fixed point relative to the channel maximum), each within the per-layer
    bars, so the reference's response to a perturbation of that size is what the product's loss can be held to."""
    from spikingjelly.activation_based import functional
    mc, sc = synth.small_config("lif")
    model = rl.build_reference_model(mc, sc, seed=0, train=True)
    B = 2
    x = synth.synth_voxels(B, 10, 96, 128)
    scales = synth.synth_drop_scales(sc["swin_depths"], B)
    blocks = [b for lyr in model.sttmultires_unet.encoders.swin3d.layers for b in lyr.swin_blocks]

    class Forced(torch.nn.Module):
        def __init__(self, s):
            super().__init__()
            self.s = s

        def forward(self, t):
            return t if self.s is None else t * self.s.view(-1, 1, 1, 1, 1)

    for b, s in zip(blocks, scales):
        b.drop_path = Forced(s)
    gt, mask = synth.synth_labels(B, 96, 128)
    sys.path.insert(0, rl.REFERENCE_ROOT)
    from loss.flow_supervised import flow_loss_supervised
    crit = flow_loss_supervised({"metrics": {"flow_scaling": 1}, "loss": {"lambda_mod": 1, "lambda_ang": 0}}, "cpu")

    def loss_with(sd):
        model.load_state_dict(sd, strict=True)
        functional.reset_net(model)
        with torch.no_grad():
            return crit(model(x)["flow"], gt, mask).item()

    sd0 = synth.synth_state_dict(model.state_dict(), 0)
    out = {"loss": loss_with(sd0), "log2_jitter": log2_jitter, "jittered": []}
    for seed in range(n_samples):
        g = torch.Generator().manual_seed(200 + seed)
        sd = {k: (v * (1 + (torch.randint(0, 2, v.shape, generator=g).float() * 2 - 1) * 2.0 ** log2_jitter)
                  if v.is_floating_point() and v.dim() >= 2 else v.clone()) for k, v in sd0.items()}
        out["jittered"].append(loss_with(sd))
        print("jitter sample", seed, out["jittered"][-1] - out["loss"], flush=True)
    return out


def golden_en4(neuron_type="lif"):
    """The shipped model/config (MS en4, window (2,9,9)) at its smallest legal size 288x384, eval."""
    from spikingjelly.activation_based import functional
    mc, sc = rl.default_config(neuron_type=neuron_type, input_size=(288, 384))
    model = rl.build_reference_model(mc, sc, seed=0, train=False)
    _load_synth(model)
    x = synth.synth_voxels(1, 10, 288, 384)
    functional.reset_net(model)
    with torch.no_grad():
        flows = model(x)["flow"]
    return {"neuron_type": neuron_type, "flows": [_summ(f) for f in flows]}


def golden_cfg4():
    """The en4 model at BASELINE.json's cfg4 shapes, eval, from the unmodified reference."""
    from spikingjelly.activation_based import functional
    out = {}
    for name, kw in synth.CFG4.items():
        mc, sc = rl.default_config("lif", **kw)
        model = rl.build_reference_model(mc, sc, seed=0, train=False)
        _load_synth(model)
        x = synth.synth_voxels(1, kw["num_bins"], *kw["input_size"])
        functional.reset_net(model)
        with torch.no_grad():
            flows = model(x)["flow"]
        out[name] = {"flows": [_summ(f) for f in flows]}
    return out


def golden_e2e_sensitivity():
    """End-to-end flows of the shipped en4 model at the BASELINE configs (cfg1/cfg2 480x640, cfg3 288x384, cfg4 shapes) from the
    unmodified reference, TOGETHER WITH the reference's own sensitivity to rounding: the same model re-evaluated with every
    F.linear / conv computed in fp64 and rounded to fp32 (tools/ref_thread_sensitivity.py, gemms_in_fp64).  A spiking net with
    hard thresholds amplifies a single threshold tie, so that half-ulp perturbation moves the reference's flow by pixels; the
    free-running end-to-end gate of the GPU tests is "within 2x of what the reference does to itself"."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from ref_thread_sensitivity import gemms_in_fp64
    from spikingjelly.activation_based import functional
    cases = {"en4_288x384": dict(kw=dict(input_size=(288, 384)), shape=(1, 10, 288, 384)),
             "en4_480x640": dict(kw=dict(input_size=(480, 640)), shape=(1, 10, 480, 640))}
    for name, kw in synth.CFG4.items():
        cases["cfg4_" + name] = dict(kw=kw, shape=(1, kw["num_bins"], *kw["input_size"]))
    out = {}
    for name, c in cases.items():
        mc, sc = rl.default_config("lif", **c["kw"])
        model = rl.build_reference_model(mc, sc, seed=0, train=False)
        _load_synth(model)
        x = synth.synth_voxels(*c["shape"])
        functional.reset_net(model)
        with torch.no_grad():
            flows = model(x)["flow"]
        functional.reset_net(model)
        with torch.no_grad(), gemms_in_fp64():
            flows64 = model(x)["flow"]
        sub = [f[..., ::8, ::8].clone() for f in flows]
        sub64 = [f[..., ::8, ::8].clone() for f in flows64]
        sens = [(a - b).pow(2).sum(1).sqrt().mean().item() for a, b in zip(sub, sub64)]
        mag = [a.pow(2).sum(1).sqrt().mean().item() for a in sub]
        out[name] = {"kw": c["kw"], "shape": c["shape"], "sub": sub, "self_sensitivity_px": sens, "flow_mag_px": mag}
        print(name, "sensitivity", [round(v, 3) for v in sens], "|flow|", [round(v, 2) for v in mag], flush=True)
    return out


def golden_sew_stage():
    """SEW family: one Spiking_Swin_BasicLayer (QK^T V attention, shifted + unshifted block, with
    SpikingPatchMerging) and the SDSA attention variant, eval and train."""
    swin, mods, sub, functional = rl.reference_modules()
    out = {}
    kw = {"num_steps": 4, "v_reset": None, "v_th": 0.3, "neuron_type": "lif", "surrogate_fun": "surrogate.ATan()",
          "tau": 2.0, "detach_reset": True, "spike_norm": "BN"}
    for variant in ("bn", "sdsa"):
        class Blk(swin.Spiking_SwinTransformerBlock3D):
            attn_module = swin.Spiking_BN_WindowAttention3D if variant == "bn" else swin.SDSA_WindowAttention3D

        class Lyr(swin.Spiking_Swin_BasicLayer):
            swin_block_type = Blk
        torch.manual_seed(0)
        lyr = Lyr(dim=64, input_resolution=(7, 10), depth=2, num_heads=2, window_size=(2, 3, 4),
                  pretrained_window_size=(0, 0, 0), mlp_ratio=4.0, version="swinv1", qk_scale=0.125,
                  drop_path=[0.0, 0.0], norm_layer="BN", downsample=swin.SpikingPatchMerging, **kw)
        lyr.load_state_dict(synth.synth_state_dict(lyr.state_dict(), seed=3))
        functional.set_step_mode(lyr, "m")
        g = torch.Generator().manual_seed(5)
        x = (torch.rand(2, 64, 4, 7, 10, generator=g) < 0.3).float() * torch.randint(1, 3, (2, 64, 4, 7, 10), generator=g)
        for train in (False, True):
            lyr.train(train)
            functional.reset_net(lyr)
            xo, xb = lyr(x.clone())
            out[f"{variant}_{'train' if train else 'eval'}"] = {"x_out": xo.detach().clone(), "x_pre": xb.detach().clone()}
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 8)
    jobs = {
        "small_lif_eval.pt": lambda: golden_small_model("lif", False),
        "small_psn_eval.pt": lambda: golden_small_model("psn", False),
        "small_lif_train.pt": lambda: golden_small_model("lif", True),
        "small_psn_train.pt": lambda: golden_small_model("psn", True),
        "small_plif_eval.pt": lambda: golden_small_model("plif", False),
        "small_plif_train.pt": lambda: golden_small_model("plif", True),
        "small_lif_train_sensitivity.pt": golden_small_train_sensitivity,
        "en4_lif_eval.pt": lambda: golden_en4("lif"),
        "sew_stage.pt": golden_sew_stage,
        "cfg4_lif_eval.pt": golden_cfg4,
        "e2e_sensitivity.pt": golden_e2e_sensitivity,
    }
    only = set(sys.argv[1:])
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        torch.save(fn(), os.path.join(OUT, name))
        print("wrote", name, os.path.getsize(os.path.join(OUT, name)) // 1024, "KiB", flush=True)


if __name__ == "__main__":
    main()
