"""TEST INFRASTRUCTURE ONLY — deterministic synthetic weights and inputs shared by the golden
generator (reference side), the CPU port and the GPU parity tests.

Weights are a pure function of (state_dict key, shape, seed) drawn from CPU generators, so the
reference model (in the build container) and the product/port (on the GPU box, where
/root/reference does not exist) get bit-identical parameters without sharing a checkpoint.
"""
import hashlib
import math
import torch

SEED_INPUT = 16146  # the reference's own smoke-test seed (Spiking_swin_transformer3D.py:1298)


def _gen(key, seed):
    h = int.from_bytes(hashlib.sha256(f"{seed}:{key}".encode()).digest()[:8], "little") & 0x7FFFFFFFFFFFFFFF
    return torch.Generator().manual_seed(h)


def synth_tensor(key, like, seed=0):
    g = _gen(key, seed)
    shape = tuple(like.shape)
    if key.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=like.dtype)
    if key.endswith("relative_position_index"):
        return like.clone()
    if key.endswith("running_mean"):
        return torch.randn(shape, generator=g) * 0.1
    if key.endswith("running_var"):
        return torch.rand(shape, generator=g) * 0.5 + 0.75
    parts = key.split(".")
    if (len(parts) >= 2 and "norm" in parts[-2]) or ".bn" in key or "_bn" in key:
        if key.endswith("weight") and len(shape) == 1:
            return torch.rand(shape, generator=g) * 0.5 + 0.75
        if key.endswith("bias") and len(shape) == 1:
            return torch.randn(shape, generator=g) * 0.1
    if key.endswith("positional_encoding"):
        return torch.randn(shape, generator=g) * 0.2
    if key.endswith("relative_position_bias_table"):
        return torch.randn(shape, generator=g) * 0.02
    if key.endswith("spiking_neuron.weight"):          # PSN [T, T]
        T = shape[0]
        return torch.eye(T) * 0.8 + torch.randn(shape, generator=g) * (0.3 / math.sqrt(T))
    if key.endswith("spiking_neuron.bias"):            # PSN [T, 1]
        return torch.full(shape, -0.1) + torch.randn(shape, generator=g) * 0.02
    if key.endswith("spiking_neuron.w"):               # PLIF scalar
        return torch.zeros(shape)
    if key.endswith("bias"):
        return torch.randn(shape, generator=g) * 0.05
    if len(shape) >= 2:                                 # Linear / Conv / ConvTranspose weights
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
    return torch.randn(shape, generator=g) * 0.1


def synth_state_dict(template, seed=0):
    """template: a state_dict (only keys/shapes/dtypes are read)."""
    return {k: synth_tensor(k, v, seed).to(v.dtype) for k, v in template.items()}


def spread_plif_w(sd, lo=-0.6, hi=0.6):
    """PLIF parameters away from their initial value: the ParametricLIFNode `w` of every site set to a different value in
    [lo, hi] (1/tau = sigmoid(w) in [0.35, 0.65]) in state_dict order, as in a trained checkpoint — synth_tensor leaves them
    at 0 (tau = 2), where a site that ignored its parameter would go unnoticed.  In place; returns sd."""
    keys = [k for k in sd if k.endswith("spiking_neuron.w")]
    for k, v in zip(keys, torch.linspace(lo, hi, max(len(keys), 1))):
        sd[k] = torch.full_like(sd[k], float(v))
    return sd


def synth_voxels(B, bins, H, W, seed=SEED_INPUT, density=0.10):
    """v = U(0,1) * [U(0,1) < density], (B, bins, 2, H, W) fp32 — post relu(+-chunk)+minmax look
    (train_flow_parallel_supervised_SNN.py:261-284)."""
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(B, bins, 2, H, W, generator=g)
    m = torch.rand(B, bins, 2, H, W, generator=g) < density
    return a * m


def synth_labels(B, H, W, seed=SEED_INPUT + 1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 2, H, W, generator=g) * 4.0, torch.ones(B, 1, H, W)


def synth_drop_scales(depths, B, drop_path_rate=0.2, seed=7):
    """Per-block DropPath factors mask/keep_prob (timm DropPath, stochastic-depth decay rule of
    Spiking_swin_transformer3D.py:1184); None for blocks with p == 0."""
    g = torch.Generator().manual_seed(seed)
    n = sum(depths)
    dpr = [x.item() for x in torch.linspace(0, drop_path_rate, n)]
    out = []
    for p in dpr:
        if p == 0.0:
            out.append(None)
        else:
            keep = 1 - p
            out.append(torch.empty(B).bernoulli_(keep, generator=g) / keep)
    return out


# the small model used by most parity tests: MS 3-encoder net, 96x128 input, window (2,3,4)
SMALL = dict(name="MS_SpikingformerFlowNet", input_size=(96, 128), window_size=(2, 3, 4), swin_depths=(2, 2, 2),
             swin_num_heads=(3, 6, 12), base_num_channels=96, num_bins=10, num_steps=10)


def small_config(neuron_type="lif", v_th=0.1, **over):
    """(model_cfg, swin_cfg) dicts shaped like the reference's YAML after combine_entries."""
    c = dict(SMALL)
    c.update(over)
    n = len(c["swin_depths"])
    model = {
        "name": c["name"], "encoding": "voxel", "norm_input": "minmax", "num_bins": c["num_bins"],
        "base_num_channels": c["base_num_channels"], "kernel_size": 3, "activations": ["relu", None],
        "final_activation": None, "mask_output": True, "norm": None, "use_upsample_conv": False,
        "spiking_neuron": {"num_steps": c["num_steps"], "v_th": v_th, "v_reset": None, "neuron_type": neuron_type,
                           "surrogate_fun": "surrogate.ATan()", "tau": 2.0, "detach_reset": True, "spike_norm": "BN"},
    }
    swin = {
        "use_arc": ["swinv1", "MS_PED_Spiking_PatchEmbed_Conv_sfn"], "state_combination": "none",
        "base_num_channels": c["base_num_channels"], "swin_depths": list(c["swin_depths"]),
        "swin_num_heads": list(c["swin_num_heads"]), "swin_out_indices": list(range(n)),
        "swin_patch_size": [1, 1, 2, 2], "window_size": list(c["window_size"]), "pretrained_window_size": [0, 0, 0],
        "mlp_ratio": 4, "input_size": list(c["input_size"]),
    }
    return model, swin


CFG4 = {
    # BASELINE.json configs[3]: MDR-shaped dt4 voxels (5 bins, window (2,8,8), 256x256 crops; configs of the MDR ymls)
    # and a larger temporal window (window_size[0] = 4)
    "t5_w288": dict(num_steps=5, num_bins=5, window_size=(2, 8, 8), input_size=(256, 256)),
    "t10_w466": dict(num_steps=10, num_bins=10, window_size=(4, 6, 6), input_size=(192, 192)),
}
