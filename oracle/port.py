"""TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference's algorithm for the hot path.

Plain PyTorch (fp32, autograd for gradients), functional style over a ``state_dict``-keyed dict
of tensors.  It follows the reference op-for-op, including the reinterpreting ``view``/``reshape``
calls and the permuted BatchNorm inputs; every function cites the reference lines it restates
(paths relative to /root/reference).  Third-party arithmetic (spikingjelly==0.0.0.0.14,
timm==0.6.13 — requirements.txt:5,17 — absent from /root/reference) is restated from
SURVEY.md Appendix A.

Pinning: the reference ships no tests or golden vectors, so this port is pinned
against outputs of the reference's own modules imported in the build container through
``oracle/standin`` (see ``oracle/make_golden.py`` -> ``tests/golden/*.pt`` and
``tests/test_oracle_golden.py``, which checks this port against those fixtures).  The stand-in itself cannot be diffed against the real
spikingjelly package offline; that residual risk is stated in DESIGN.md.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
import math
from functools import reduce
from operator import mul

import torch
import torch.nn.functional as F

# The reference's explicit .float() casts (Spiking_swin_transformer3D.py:670,675,710) are no-ops in its
# fp32 path.  Tests that measure the fp32 path's own rounding sensitivity run this port in float64 and
# set KEEP_DTYPE so those casts do not truncate.
KEEP_DTYPE = False


def _f(x):
    return x if KEEP_DTYPE else x.float()


# ---------------------------------------------------------------------------------------------
# spikingjelly semantics
# ---------------------------------------------------------------------------------------------
class _ATan(torch.autograd.Function):
    """surrogate.ATan: forward heaviside(x) = (x >= 0); backward g * alpha/2 / (1 + (pi/2*alpha*x)^2)."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.save_for_backward(x)
        ctx.alpha = alpha
        return (x >= 0).to(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ctx.alpha / 2 / (1 + (math.pi / 2 * ctx.alpha * x).pow(2)) * g, None


def heaviside_atan(x, alpha=2.0):
    return _ATan.apply(x, alpha)


class NeuronSpec:
    """Mirror of the kwargs of Spiking_neuron (models/STSwinNet_SNN/Spiking_modules.py:27-37)."""

    def __init__(self, num_steps, neuron_type="lif", v_th=1.0, v_reset=0.0, tau=2.0, detach_reset=True,
                 sg_alpha=2.0, spike_norm="BN", surrogate_fun=None):
        self.num_steps, self.neuron_type, self.v_th, self.v_reset = num_steps, neuron_type, v_th, v_reset
        self.tau, self.detach_reset, self.sg_alpha, self.spike_norm = tau, detach_reset, sg_alpha, spike_norm

    def with_steps(self, T):
        s = NeuronSpec.__new__(NeuronSpec)
        s.__dict__.update(self.__dict__)
        s.num_steps = T
        return s


def lif_multistep(x_seq, spec, kind="lif", plif_w=None, record=None):
    """BaseNode.multi_step_forward over x_seq[t] with LIF / IF / PLIF charge, ATan fire, soft or hard
    reset (spikingjelly neuron.py; reached from Spiking_modules.py:41-77,98-99).  v starts at 0
    (v_reset None) or v_reset: the scripts call functional.reset_net before every sample
    (train_flow_parallel_supervised_SNN.py:238, eval_DSEC_flow_SNN.py:155)."""
    v_reset = spec.v_reset
    v = torch.zeros_like(x_seq[0]) if v_reset is None else torch.full_like(x_seq[0], float(v_reset))
    out, hs = [], []
    for t in range(x_seq.shape[0]):
        x = x_seq[t]
        if kind == "if":
            v = v + x
        elif kind == "plif":
            k = plif_w.sigmoid()
            v = v + (x - v) * k if (v_reset is None or v_reset == 0.0) else v + (x - (v - v_reset)) * k
        else:
            v = v + (x - v) / spec.tau if (v_reset is None or v_reset == 0.0) else v + (x - (v - v_reset)) / spec.tau
        if record is not None:
            hs.append(v)
        s = heaviside_atan(v - spec.v_th, spec.sg_alpha)
        sd = s.detach() if spec.detach_reset else s
        v = v - sd * spec.v_th if v_reset is None else (1.0 - sd) * v + sd * v_reset
        out.append(s)
    if record is not None:
        record.append(torch.stack(hs))
    return torch.stack(out)


def psn_forward(x_seq, weight, bias, sg_alpha=2.0, record=None):
    """PSN.forward (models/STSwinNet_SNN/Spiking_submodules.py:207-211)."""
    h = torch.addmm(bias, weight, x_seq.flatten(1))
    if record is not None:
        record.append(h.view(x_seq.shape))
    return heaviside_atan(h, sg_alpha).view(x_seq.shape)


def spiking_neuron(x_seq, P, prefix, spec, record=None):
    """Spiking_neuron.forward (Spiking_modules.py:98-99) dispatching on neuron_type (:40-96)."""
    nt = spec.neuron_type
    if nt == "psn":
        return psn_forward(x_seq, P[prefix + ".spiking_neuron.weight"], P[prefix + ".spiking_neuron.bias"],
                           spec.sg_alpha, record)
    if nt == "plif":
        return lif_multistep(x_seq, spec, "plif", P[prefix + ".spiking_neuron.w"], record)
    if nt in ("lif", "if"):
        return lif_multistep(x_seq, spec, nt, None, record)
    raise NotImplementedError(nt)


class BNMode:
    """training flag + optional dict collecting the updated running statistics."""

    def __init__(self, training=False, momentum=0.1, eps=1e-5):
        self.training, self.momentum, self.eps = training, momentum, eps


def batchnorm_seq(x, P, prefix, mode):
    """sj_layer.BatchNorm2d in 'm' mode: BN2d on flatten(0,1) of [T,B,C,H,W] (stats over T*B*H*W),
    reached through SpikingNormLayer.forward (Spiking_modules.py:133-146).  Running stats in P are
    updated in place when training, like torch."""
    y = F.batch_norm(x.flatten(0, 1), P[prefix + ".running_mean"], P[prefix + ".running_var"],
                     P[prefix + ".weight"], P[prefix + ".bias"], mode.training, mode.momentum, mode.eps)
    if mode.training:
        P[prefix + ".num_batches_tracked"] += 1
    return y.view(x.shape)


def batchnorm_4d(x, P, prefix, mode):
    y = F.batch_norm(x, P[prefix + ".running_mean"], P[prefix + ".running_var"], P[prefix + ".weight"],
                     P[prefix + ".bias"], mode.training, mode.momentum, mode.eps)
    if mode.training:
        P[prefix + ".num_batches_tracked"] += 1
    return y


def conv_seq(x, w, b=None, stride=1, padding=0):
    """sj_layer.Conv2d in 'm' mode (functional.seq_to_ann_forward)."""
    y = F.conv2d(x.flatten(0, 1), w, b, stride=stride, padding=padding)
    return y.view(x.shape[0], x.shape[1], *y.shape[1:])


def deconv_seq(x, w, b=None, stride=2, padding=1, output_padding=1):
    y = F.conv_transpose2d(x.flatten(0, 1), w, b, stride=stride, padding=padding, output_padding=output_padding)
    return y.view(x.shape[0], x.shape[1], *y.shape[1:])


# ---------------------------------------------------------------------------------------------
# window helpers (models/STSwinNet/swin_transformer3D_v2.py:37-81, Spiking_swin_transformer3D.py:100-113,980-993)
# ---------------------------------------------------------------------------------------------
def get_window_size(x_size, window_size, shift_size=None):
    """swin_transformer3D_v2.py:68-81"""
    ws = list(window_size)
    ss = list(shift_size) if shift_size is not None else None
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            ws[i] = x_size[i]
            if ss is not None:
                ss[i] = 0
    return tuple(ws) if ss is None else (tuple(ws), tuple(ss))


def window_partition(x, ws):
    """swin_transformer3D_v2.py:37-49"""
    B, D, H, W, C = x.shape
    x = x.view(B, D // ws[0], ws[0], H // ws[1], ws[1], W // ws[2], ws[2], C)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(-1, reduce(mul, ws), C)


def window_partition_v2(x, ws):
    """Spiking_swin_transformer3D.py:100-113 — note the reinterpreting .view at the end."""
    B, D, H, W, C = x.shape
    x = x.view(B, D // ws[0], ws[0], H // ws[1], ws[1], W // ws[2], ws[2], C)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(ws[0], -1, ws[1], ws[2], C)


def window_reverse(windows, ws, B, D, H, W):
    """swin_transformer3D_v2.py:52-65"""
    x = windows.view(B, D // ws[0], H // ws[1], W // ws[2], ws[0], ws[1], ws[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(B, D, H, W, -1)


def compute_mask(D, H, W, ws, ss):
    """Spiking_swin_transformer3D.py:980-993"""
    img_mask = torch.zeros((1, D, H, W, 1))
    cnt = 0
    for d in slice(-ws[0]), slice(-ws[0], -ss[0]), slice(-ss[0], None):
        for h in slice(-ws[1]), slice(-ws[1], -ss[1]), slice(-ss[1], None):
            for w in slice(-ws[2]), slice(-ws[2], -ss[2]), slice(-ss[2], None):
                img_mask[:, d, h, w, :] = cnt
                cnt += 1
    mw = window_partition(img_mask, ws).squeeze(-1)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0))


def relative_position_index(ws):
    """Spiking_swin_transformer3D.py:251-265"""
    coords = torch.stack(torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), torch.arange(ws[2]), indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws[0] - 1
    rel[:, :, 1] += ws[1] - 1
    rel[:, :, 2] += ws[2] - 1
    rel[:, :, 0] *= (2 * ws[1] - 1) * (2 * ws[2] - 1)
    rel[:, :, 1] *= (2 * ws[2] - 1)
    return rel.sum(-1)


def _bn_perm(x, P, prefix, mode):
    """bn(x.permute(0,1,4,2,3)).permute(0,1,3,4,2) on (T,B_,H,W,C) — e.g. :153,:310,:673."""
    return batchnorm_seq(x.permute(0, 1, 4, 2, 3), P, prefix + ".norm_layer", mode).permute(0, 1, 3, 4, 2)


# ---------------------------------------------------------------------------------------------
# attention variants
# ---------------------------------------------------------------------------------------------
def qk_window_attention(x, P, pre, num_heads, spec, mode, rec=None):
    """Spiking_QK_WindowAttention3D.forward (Spiking_swin_transformer3D.py:661-717).
    x: (T=wd, B_, wh, ww, C) from window_partition_v2.  Returns (x (B_, N, C), attn-score spikes)."""
    T, B_, H, W, C = x.shape
    sp = spec.with_steps(T)
    x = spiking_neuron(_f(x), P, pre + ".proj_sn", sp, rec)
    q = F.linear(x, P[pre + ".linear_q.weight"])
    q = _bn_perm(q, P, pre + ".bn_q", mode)
    q = spiking_neuron(q, P, pre + ".sn_q", sp, rec)
    k = _f(F.linear(x, P[pre + ".linear_k.weight"]))
    k = _bn_perm(k, P, pre + ".bn_k", mode)
    k = k + P[pre + ".positional_encoding"].reshape(T, 1, H, W, C)
    k = spiking_neuron(k, P, pre + ".sn_k", sp, rec)
    hd = C // num_heads
    q, k = q.reshape(T, B_, num_heads, -1, hd), k.reshape(B_, num_heads, -1, hd)
    N = k.shape[2]
    att_token = q.sum(dim=-1, keepdim=True)
    att_token = spiking_neuron(att_token, P, pre + ".sn2_q", sp, rec)
    attn = k.mul(att_token.reshape(B_, num_heads, -1, 1))
    x = attn.reshape(B_, num_heads, T, H, W, hd)
    x = _f(x.permute(2, 0, 3, 4, 1, 5).reshape(T, B_, H, W, C))
    gate = x
    x = F.linear(x, P[pre + ".proj.weight"], P[pre + ".proj.bias"])
    x = _bn_perm(x, P, pre + ".proj_bn", mode)
    return x.reshape(B_, N, C), gate


def qktv_window_attention(x, P, pre, num_heads, spec, mode, mask, qk_scale, variant="bn", rec=None):
    """Spiking_BN_WindowAttention3D.forward (:300-370, variant 'bn') and
    SDSA_WindowAttention3D.forward (:416-492, variant 'sdsa'); swinv1 only (swinv2 is dead)."""
    T, B_, H, W, C = x.shape
    sp = spec.with_steps(T)
    if variant == "sdsa":
        x = spiking_neuron(x, P, pre + ".proj_sn", sp, rec)

    def branch(name):
        y = F.linear(x, P[pre + f".linear_{name}.weight"])
        y = _bn_perm(y, P, pre + f".bn_{name}", mode)
        return spiking_neuron(y, P, pre + f".sn_{name}", sp, rec)

    q, k, v = branch("q"), branch("k"), branch("v")
    hd = C // num_heads
    q, k, v = (t.reshape(B_, num_heads, -1, hd) for t in (q, k, v))
    N = q.shape[2]
    scale = 1 if spec.neuron_type in ("psn", "glif") else (qk_scale or hd ** -0.5)
    attn = (q * scale) @ k.transpose(-2, -1)                                       # VanillaAttention :18-29
    idx = P[pre + ".relative_position_index"][:N, :N].reshape(-1)
    bias = P[pre + ".relative_position_bias_table"][idx].reshape(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, num_heads, N, N)
    x = (attn @ v).reshape(B_, num_heads, T, H, W, hd)
    x = x.permute(2, 0, 3, 4, 1, 5).reshape(T, B_, H, W, C)
    x = F.linear(x, P[pre + ".proj.weight"], P[pre + ".proj.bias"])
    x = _bn_perm(x, P, pre + ".proj_bn", mode)
    if variant == "bn":
        x = spiking_neuron(x, P, pre + ".proj_sn", sp, rec)
    return x.reshape(B_, N, C), attn


# ---------------------------------------------------------------------------------------------
# MLP, block, merging, stage, backbone
# ---------------------------------------------------------------------------------------------
def ms_mlp(x, P, pre, spec, mode, rec=None):
    """MS_Spiking_Mlp.forward (:164-181): sn1 -> fc1 -> bn1 -> sn2 -> fc2 -> bn2; x is (D,B,H,W,C)."""
    x = spiking_neuron(x, P, pre + ".sn1", spec, rec)
    x = F.linear(x, P[pre + ".fc1.weight"])
    x = _bn_perm(x, P, pre + ".bn1", mode)
    x = spiking_neuron(x, P, pre + ".sn2", spec, rec)
    x = F.linear(x, P[pre + ".fc2.weight"])
    return _bn_perm(x, P, pre + ".bn2", mode)


def sew_mlp(x, P, pre, spec, mode, rec=None):
    """Spiking_Mlp.forward (:147-162): fc1 -> bn1 -> sn1 -> fc2 -> bn2 -> sn2."""
    x = F.linear(x, P[pre + ".fc1.weight"])
    x = _bn_perm(x, P, pre + ".bn1", mode)
    x = spiking_neuron(x, P, pre + ".sn1", spec, rec)
    x = F.linear(x, P[pre + ".fc2.weight"])
    x = _bn_perm(x, P, pre + ".bn2", mode)
    return spiking_neuron(x, P, pre + ".sn2", spec, rec)


class SwinCfg:
    def __init__(self, window_size=(2, 9, 9), depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), embed_dim=96,
                 mlp_ratio=4.0, qk_scale=0.125, family="ms", attn="qk", drop_path_rate=0.2):
        self.window_size, self.depths, self.num_heads = tuple(window_size), tuple(depths), tuple(num_heads)
        self.embed_dim, self.mlp_ratio, self.qk_scale = embed_dim, mlp_ratio, qk_scale
        self.family, self.attn, self.drop_path_rate = family, attn, drop_path_rate


def swin_block(x, P, pre, cfg, num_heads, shift_size, mask_matrix, spec, mode, drop_scale=None, rec=None):
    """(MS_)Spiking_SwinTransformerBlock3D.forward + SSA (:781-847).
    drop_scale: None or (B,) tensor = DropPath mask / keep_prob for this block (timm DropPath)."""
    B, D, H, W, C = x.shape
    shortcut = x
    ws, ss = get_window_size((D, H, W), cfg.window_size, shift_size)
    pad_d1 = (ws[0] - D % ws[0]) % ws[0]
    pad_b = (ws[1] - H % ws[1]) % ws[1]
    pad_r = (ws[2] - W % ws[2]) % ws[2]
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b, 0, pad_d1))
    _, Dp, Hp, Wp, _ = x.shape
    if any(i > 0 for i in ss):
        shifted = torch.roll(x, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
        attn_mask = mask_matrix
    else:
        shifted, attn_mask = x, None
    xw = window_partition_v2(shifted, ws)
    if cfg.attn == "qk":
        aw, _ = qk_window_attention(xw, P, pre + ".attn", num_heads, spec, mode, rec)
    else:
        aw, _ = qktv_window_attention(xw, P, pre + ".attn", num_heads, spec, mode, attn_mask, cfg.qk_scale,
                                      cfg.attn, rec)
    aw = aw.view(-1, *(ws + (C,)))
    shifted = window_reverse(aw, ws, B, Dp, Hp, Wp)
    x = torch.roll(shifted, shifts=ss, dims=(1, 2, 3)) if any(i > 0 for i in ss) else shifted
    if pad_d1 > 0 or pad_r > 0 or pad_b > 0:
        x = x[:, :D, :H, :W, :].contiguous()
    if drop_scale is not None:
        x = x * drop_scale.view(B, 1, 1, 1, 1)
    x = x + shortcut
    mlp = ms_mlp if cfg.family == "ms" else sew_mlp
    return mlp(x.permute(1, 0, 2, 3, 4), P, pre + ".mlp", spec, mode, rec).permute(1, 0, 2, 3, 4) + x


def patch_merging(x, P, pre, cfg, spec, mode, rec=None):
    """MS_SpikingPatchMerging.forward (:953-974) / SpikingPatchMerging.forward (:914-935).
    Returns (B, D, H/2, W/2, 2C)."""
    B, D, H, W, C = x.shape
    if (H % 2 == 1) or (W % 2 == 1):
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, :, 0::2, 0::2, :], x[:, :, 1::2, 0::2, :], x[:, :, 0::2, 1::2, :], x[:, :, 1::2, 1::2, :]], -1)
    if cfg.family == "ms":
        x = spiking_neuron(x.permute(1, 0, 2, 3, 4), P, pre + ".sn", spec, rec)
        x = F.linear(x, P[pre + ".reduction.weight"])
        return batchnorm_seq(x.permute(0, 1, 4, 2, 3), P, pre + ".norm.norm_layer", mode).permute(1, 0, 3, 4, 2)
    x = F.linear(x.permute(1, 0, 2, 3, 4), P[pre + ".reduction.weight"])
    x = batchnorm_seq(x.permute(0, 1, 4, 2, 3), P, pre + ".norm.norm_layer", mode).permute(0, 1, 3, 4, 2)
    return spiking_neuron(x, P, pre + ".sn", spec, rec).permute(1, 0, 2, 3, 4)


def basic_layer(x, P, pre, cfg, i_layer, spec, mode, drop_scales=None, rec=None):
    """Spiking_Swin_BasicLayer.forward (:1065-1088) on (B, D, H, W, C) channels-last input.
    Returns (x_out (B,D,H',W',C'), x_before_merging)."""
    B, D, H, W, C = x.shape
    shift_full = tuple(i // 2 for i in cfg.window_size)
    ws, ss = get_window_size((D, H, W), cfg.window_size, shift_full)
    Dp, Hp, Wp = (int(math.ceil(n / w)) * w for n, w in zip((D, H, W), ws))
    mask = compute_mask(Dp, Hp, Wp, ws, ss) if cfg.attn != "qk" else None
    for k in range(cfg.depths[i_layer]):
        shift = (0, 0, 0) if k % 2 == 0 else shift_full
        ds = None if drop_scales is None else drop_scales[k]
        x = swin_block(x, P, f"{pre}.swin_blocks.{k}", cfg, cfg.num_heads[i_layer], shift, mask, spec, mode, ds, rec)
    if i_layer < len(cfg.depths) - 1:
        return patch_merging(x, P, pre + ".downsample", cfg, spec, mode, rec), x
    return x, x


def swin_stages(x, P, pre, cfg, spec, mode, drop_scales=None, rec=None):
    """layers loop of Spiking_SwinTransformer3D_v2.forward (:1230-1246); x: (B, D, H, W, C).
    Returns the per-stage features (before merging), each (B, D, Hi, Wi, Ci)."""
    outs, blk = [], 0
    for i in range(len(cfg.depths)):
        ds = None if drop_scales is None else drop_scales[blk:blk + cfg.depths[i]]
        blk += cfg.depths[i]
        x, out = basic_layer(x, P, f"{pre}.layers.{i}", cfg, i, spec, mode, ds, rec)
        outs.append(out)
    return outs


# ---------------------------------------------------------------------------------------------
# the rest of MS_SpikingformerFlowNet(_en4): patch embed, res blocks, decoders (callers of the path)
# ---------------------------------------------------------------------------------------------
def regroup_events(x, num_bins, num_steps):
    """MS_PED_Spiking_PatchEmbed_Conv_sfn.forward bins->steps regroup (Spiking_modules.py:1772-1784)."""
    if x.size(1) > num_bins:
        x = x[:, :num_bins]
    num_ch = num_bins * 2 // num_steps
    ev = x.permute(0, 2, 3, 4, 1)
    new = torch.zeros(ev.size(0), num_ch, ev.size(2), ev.size(3), num_steps, dtype=ev.dtype, device=ev.device)
    for i in range(num_ch):
        s, e = i // 2 * num_steps, (i // 2 + 1) * num_steps
        new[:, i] = ev[:, i % 2, :, :, s:e]
    return new.permute(4, 0, 1, 2, 3)


def ms_resblock(x, P, pre, spec, mode, rec=None):
    """MS_ResBlock.forward (Spiking_modules.py:906-933): sn1->conv1->norm1->sn2->conv2->norm2 + identity."""
    idt = x
    x = spiking_neuron(x, P, pre + ".sn1", spec, rec)
    x = conv_seq(x, P[pre + ".conv1.0.weight"], None, 1, 1)
    x = batchnorm_seq(x, P, pre + ".norm1.norm_layer", mode)
    x = spiking_neuron(x, P, pre + ".sn2", spec, rec)
    x = conv_seq(x, P[pre + ".conv2.0.weight"], None, 1, 1)
    x = batchnorm_seq(x, P, pre + ".norm2.norm_layer", mode)
    return x + idt


def patch_embed_ms_ped(x, P, pre, spec, mode, num_bins, rec=None):
    """MS_PED_Spiking_PatchEmbed_Conv_sfn.forward (Spiking_modules.py:1770-1790)."""
    x = regroup_events(x, num_bins, spec.num_steps)
    # head: SpikingConvEncoderLayer (:291-296) conv -> BN -> sn
    x = conv_seq(x, P[pre + ".head.conv.0.weight"], None, 1, 1)
    x = batchnorm_seq(x, P, pre + ".head.norm_layer.norm_layer", mode)
    x = spiking_neuron(x, P, pre + ".head.sn", spec, rec)
    # conv: MS_SpikingConvEncoderLayer first_layer=True (:339-347) conv s2 -> BN
    x = conv_seq(x, P[pre + ".conv.conv.0.weight"], None, 2, 1)
    x = batchnorm_seq(x, P, pre + ".conv.norm_layer.norm_layer", mode)
    for i in range(2):
        x = ms_resblock(x, P, f"{pre}.residual_encoding.resblocks.{i}", spec, mode, rec)
    # proj: SpikingPEDLayer.forward (:816-825)
    T, B, C, H, W = x.shape
    x_res = F.conv2d(x.flatten(0, 1), P[pre + ".proj.conv_res.weight"], None, stride=2, padding=0)
    y = spiking_neuron(x, P, pre + ".proj.sn", spec, rec)
    y = F.conv2d(y.flatten(0, 1), P[pre + ".proj.conv.weight"], None, stride=2, padding=1)
    y = batchnorm_4d(y, P, pre + ".proj.norm_layer", mode)
    y = y + x_res
    return y.reshape(T, B, -1, y.shape[-2], y.shape[-1]).contiguous()


def skip_concat(x1, x2, dim):
    """models/model_util.py:13-18"""
    dY, dX = x2.size(-2) - x1.size(-2), x2.size(-1) - x1.size(-1)
    x1 = F.pad(x1, (dX // 2, dX - dX // 2, dY // 2, dY - dY // 2))
    return torch.cat([x1, x2], dim=dim)


class FlowNetCfg:
    def __init__(self, swin: SwinCfg, num_bins=10, num_steps=10):
        self.swin, self.num_bins, self.num_steps = swin, num_bins, num_steps


def ms_flownet_forward(x, P, cfg, spec, mode, drop_scales=None, rec=None, feats=None):
    """MS_SpikingformerFlowNet(_en4).forward (Spiking_STSwinNet.py:278-305) =
    MS_Spikingformer_MultiResUNet.forward (:161-182) + flow accumulation.  x: (B, bins, 2, H, W)."""
    Hin, Win = x.shape[-2], x.shape[-1]
    u = "sttmultires_unet"
    pe = patch_embed_ms_ped(x, P, f"{u}.encoders.swin3d.patch_embed", spec, mode, cfg.num_bins, rec)  # (T,B,C,H,W)
    xs = pe.permute(1, 0, 3, 4, 2).contiguous()                                                   # (B,D,H,W,C)
    outs = swin_stages(xs, P, f"{u}.encoders.swin3d", cfg.swin, spec, mode, drop_scales, rec)
    blocks = [o.permute(1, 0, 4, 2, 3) for o in outs]                                             # (T,B,C,H,W)
    if feats is not None:
        feats.extend(outs)
    n = len(blocks)
    x = blocks[-1]
    for i in range(2):
        x = ms_resblock(x, P, f"{u}.resblocks.{i}", spec, mode, rec)
    preds = []
    for i in range(n):
        x = skip_concat(x, blocks[n - i - 1], dim=2)
        if i > 0:
            x = skip_concat(preds[-1], x, dim=2)
        # MS_SpikingTransposeDecoderLayer.forward (Spiking_modules.py:467-474): sn -> deconv -> BN
        x = spiking_neuron(x, P, f"{u}.decoders.{i}.sn", spec, rec)
        x = deconv_seq(x, P[f"{u}.decoders.{i}.deconv.0.weight"], None, 2, 1, 1)
        x = batchnorm_seq(x, P, f"{u}.decoders.{i}.norm_layer.norm_layer", mode)
        # MS_SpikingPredLayer.forward (:643-647): sn -> 1x1 conv (+bias)
        p = spiking_neuron(x, P, f"{u}.preds.{i}.sn", spec, rec)
        p = conv_seq(p, P[f"{u}.preds.{i}.conv.0.weight"], P[f"{u}.preds.{i}.conv.0.bias"], 1, 0)
        preds.append(p)
    flows = []
    for f in preds:
        f = torch.sum(f, dim=0)
        flows.append(F.interpolate(f, scale_factor=(Hin / f.shape[-2], Win / f.shape[-1])))
    return flows


def flow_loss(pred_list, gt, mask):
    """flow_loss_supervised.forward with gamma None, lambda_mod 1, flow_scaling 1 (loss/flow_supervised.py:14-31,81-105)."""
    nv = torch.sum(mask)
    cur = 0.0
    for pred in pred_list:
        err = torch.sqrt((pred - gt).pow(2).sum(1) + 1e-8).view(pred.shape[0], -1) * mask.reshape(pred.shape[0], -1)
        cur = cur + torch.sum(err, dim=1) / (nv + 1e-9)
    return torch.mean(cur / len(pred_list))


def params_from_state_dict(sd, requires_grad=False, device="cpu"):
    """Leaf fp32 copies of a state_dict (buffers stay plain tensors)."""
    P = {}
    for k, v in sd.items():
        t = v.detach().to(device).clone()
        if requires_grad and t.is_floating_point() and not any(
                s in k for s in ("running_mean", "running_var", "relative_position_index")):
            t.requires_grad_(True)
        P[k] = t
    return P
