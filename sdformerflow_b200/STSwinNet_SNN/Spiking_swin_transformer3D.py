"""Spiking 3D shifted-window Swin encoder on the sm_100a kernels (host-side mirror of reference
models/STSwinNet_SNN/Spiking_swin_transformer3D.py: same classes, ctor kwargs, state_dict keys).

Data layout: the residual stream stays ONE contiguous channels-last tensor (B, D, H, W, C) for
a whole stage.  The reference's permutes / rearranges / window_partition / roll / pad / crop /
window_reverse copies are all folded into kernel indexing:
  * neurons over real time read (B, D, ...) with a time stride (no x.permute(1,0,2,3,4) copy),
  * BatchNorm on "permuted views" is BN over channels-last rows (statistics are order-free),
  * the window machinery is one cached int32 table per (shape, shift) (ops.WindowGeom).
Linear layers on spike operands run on the library's own tcgen05 + TMA GEMM engine (ops.spike_linear on ops.Spikes:
1-byte spikes, kind::i8 forward with the BatchNorm statistics from the epilogue, TF32 gradients); real-valued operands
(SEW proj input) still go through cuBLAS.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..sj import layer as sj_layer
from ..sj import surrogate, neuron  # noqa: F401
from .. import gemm, ops
from .Spiking_modules import *  # noqa: F401,F403
from .Spiking_modules import Spiking_neuron, SpikingNormLayer, MS_PED_Spiking_PatchEmbed_Conv_sfn, bn_training  # noqa: F401


def get_window_size(x_size, window_size, shift_size=None):
    """Clamp window (and shift) on axes not larger than the window (reference swin_transformer3D_v2.py:68-81)."""
    ws = list(window_size)
    ss = None if shift_size is None else list(shift_size)
    for i, n in enumerate(x_size):
        if n <= window_size[i]:
            ws[i] = n
            if ss is not None:
                ss[i] = 0
    return tuple(ws) if ss is None else (tuple(ws), tuple(ss))


class DropPath(nn.Module):
    """Stochastic depth with timm's semantics and RNG consumption (one bernoulli_ draw of B values per
    call in train mode).  `scale(x)` returns the per-sample factor mask/keep_prob that the window
    scatter kernel applies; `forced` lets a test inject the oracle's mask."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep
        self.forced = None

    def scale(self, x):
        if self.forced is not None:
            return self.forced.to(device=x.device, dtype=torch.float32).contiguous()
        if self.drop_prob == 0.0 or not self.training:
            return None
        keep = 1 - self.drop_prob
        r = x.new_empty((x.shape[0],)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            r.div_(keep)
        return r

    def forward(self, x):
        s = self.scale(x)
        return x if s is None else x * s.view(-1, *([1] * (x.ndim - 1)))

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


def _is_bn(norm):
    return norm in ["BN", "BNTT", "tdBN", "IN"]


def _bn_of(sn_layer):
    """The nn.BatchNorm2d inside a SpikingNormLayer; only plain 'BN' is fused."""
    if not sn_layer.is_batchnorm:
        raise NotImplementedError(f"spike norm {sn_layer.norm!r}: only 'BN' is built on the B200 hot path")
    return sn_layer.norm_layer


# ---------------------------------------------------------------------------------------------
# MLP
# ---------------------------------------------------------------------------------------------
class Spiking_Mlp(nn.Module):
    """SEW MLP: fc1 -> bn1 -> sn1 -> fc2 -> bn2 -> sn2 (reference :115-162)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, norm_layer="BN", act_layer=nn.GELU,
                 drop=0.0, **spiking_kwargs):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.norm_layer = norm_layer
        self.fc1 = sj_layer.Linear(in_features, hidden_features, bias=False)
        if _is_bn(norm_layer):
            self.bn1 = SpikingNormLayer(hidden_features, spiking_kwargs["num_steps"], spiking_kwargs["spike_norm"],
                                        v_th=spiking_kwargs["v_th"])
        self.sn1 = Spiking_neuron(**spiking_kwargs)
        self.fc2 = sj_layer.Linear(hidden_features, out_features, bias=False)
        if _is_bn(norm_layer):
            self.bn2 = SpikingNormLayer(out_features, spiking_kwargs["num_steps"], spiking_kwargs["spike_norm"],
                                        v_th=spiking_kwargs["v_th"])
        self.sn2 = Spiking_neuron(**spiking_kwargs)
        if norm_layer in ["LN", "GN"]:
            raise NotImplementedError("LN/GN spike norm is not built on the B200 hot path (shipped configs use BN)")
        self.drop1 = sj_layer.Dropout(drop)
        self.drop2 = sj_layer.Dropout(drop)
        if drop != 0.0:
            raise NotImplementedError("MLP dropout > 0 is not built (reference passes drop_rate=0, Spiking_STSwinNet.py:63)")

    def _bn_sn(self, h, bn, sn, time_dim, partials=None, u8=False):
        sn.mark()
        return ops.bn_neuron(h, _bn_of(bn), sn.cfg(), time_dim, psn=sn.spiking_neuron if sn.is_psn else None,
                             plif_w=sn.plif_w(), partials=partials, u8=u8 and ops.spike_gemm_on())

    def forward(self, x, time_dim=0):
        """x: [T, B, H, W, C] (time_dim 0, the reference's call) or (B, D, H, W, C) (time_dim 1)."""
        h = ops.spike_linear(x, self.fc1.weight)           # SEW stream: spikes + residual adds = small integers (fp32)
        s = self._bn_sn(h, self.bn1, self.sn1, time_dim, u8=True)
        h, part = ops.spike_linear(s, self.fc2.weight, stats=bn_training(self.bn2))
        return self._bn_sn(h, self.bn2, self.sn2, time_dim, part)

    def fused(self, x):
        """block tail on the (B, D, H, W, C) stream: mlp(x) + x (reference :845, cnf ADD)."""
        return self.forward(x, time_dim=1) + x


class MS_Spiking_Mlp(Spiking_Mlp):
    """MS MLP: sn1 -> fc1 -> bn1 -> sn2 -> fc2 -> bn2 (reference :164-181)."""

    def forward(self, x, time_dim=0, res=None):
        s = self.sn1(x, time_dim, u8=True)
        h, part = ops.spike_linear(s, self.fc1.weight, stats=bn_training(self.bn1))
        s = self._bn_sn(h, self.bn1, self.sn2, time_dim, part, u8=True)
        h, part = ops.spike_linear(s, self.fc2.weight, stats=bn_training(self.bn2))
        return ops.bn_residual(h, _bn_of(self.bn2), res, partials=part)

    def fused(self, x):
        return self.forward(x, time_dim=1, res=x)


# ---------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------
class _WindowAttentionBase(nn.Module):
    def __init__(self, dim, window_size, pretrained_window_size, num_heads, version, norm, spiking_kwargs):
        super().__init__()
        self.dim, self.window_size = dim, tuple(window_size)
        self.pretrained_window_size, self.num_heads = pretrained_window_size, num_heads
        self.version, self.norm_layer = version, norm
        if version != "swinv1":
            # the reference's swinv2 SNN branch reads self.Ham_attn, which is never constructed (:286,:336)
            raise NotImplementedError("only use_arc[0] == 'swinv1' works for the spiking Swin (also in the reference)")
        if not _is_bn(norm):
            raise NotImplementedError("spiking window attention is built for BN spike norm only")
        if dim != num_heads * 32:
            raise NotImplementedError("head_dim must be 32 (96/3 = 192/6 = 384/12 = 768/24 in every reference model)")
        spiking_kwargs["num_steps"] = self.window_size[0]   # neurons inside attention run over the window depth

    def _geom_args(self, x):
        T, B_, H, W, C = x.shape
        return T, B_, H * W, C


class Spiking_QK_WindowAttention3D(_WindowAttentionBase):
    """QK token-gate window attention (reference :605-717).  No N x N matrix, no V, mask ignored."""

    def __init__(self, dim, window_size, pretrained_window_size, num_heads, version="swinv1", qkv_bias=False,
                 qk_scale=None, attn_drop=0.0, proj_drop=0.0, norm=None, **spiking_kwargs):
        super().__init__(dim, window_size, pretrained_window_size, num_heads, version, norm, spiking_kwargs)
        head_dim = dim // num_heads
        self.scale = 1 if spiking_kwargs["neuron_type"] in ["psn", "glif"] else (qk_scale or head_dim ** -0.5)
        ws = self.window_size
        self.positional_encoding = nn.Parameter(torch.zeros(size=(1, num_heads, ws[0] * ws[1] * ws[2], head_dim)))
        self.linear_q = sj_layer.Linear(dim, dim, bias=False)
        self.bn_q = SpikingNormLayer(dim, ws[0], norm, spiking_kwargs["v_th"])
        self.sn_q = Spiking_neuron(**spiking_kwargs)
        self.linear_k = sj_layer.Linear(dim, dim, bias=False)
        self.bn_k = SpikingNormLayer(dim, ws[0], norm, spiking_kwargs["v_th"])
        self.sn_k = Spiking_neuron(**spiking_kwargs)
        self.sn2_q = Spiking_neuron(**spiking_kwargs)
        self.attn_sn = Spiking_neuron(**spiking_kwargs)
        self.attn_drop = sj_layer.Dropout(attn_drop)
        self.proj = sj_layer.Linear(dim, dim)
        self.proj_bn = SpikingNormLayer(dim, ws[0], norm, spiking_kwargs["v_th"])
        self.proj_sn = Spiking_neuron(**spiking_kwargs)
        self.proj_drop = sj_layer.Dropout(proj_drop)

    def _wqk(self):
        """[W_q; W_k] so that q_pre | k_pre come from ONE GEMM.  In eval the concatenation (and its packed digit planes)
        is cached per parameter version; in training autograd needs the cat node every step."""
        wq, wk = self.linear_q.weight, self.linear_k.weight
        if torch.is_grad_enabled() and (wq.requires_grad or wk.requires_grad):
            return torch.cat([wq, wk], 0)
        key = (wq.data_ptr(), wq._version, wk.data_ptr(), wk._version, gemm.weights_epoch())
        hit = self.__dict__.get("_wqk_cache")
        if hit is None or hit[0] != key or torch.cuda.is_current_stream_capturing():
            w = torch.cat([wq.detach(), wk.detach()], 0)
            w._sdf_cacheable = True
            hit = self.__dict__["_wqk_cache"] = (key, w)
        return hit[1]

    def _core(self, s, wd, M, P):
        """s: input spikes (after proj_sn), ops.Spikes or fp32 [wd*M*P, C] -> (gate, proj output rows [wd*M*P, C]
        before proj_bn, BN partial sums of those rows or None)."""
        C, nH = self.dim, self.num_heads
        if not self.sn_q.fusable:
            return self._core_generic(s, wd, M, P)
        for sn in (self.sn_q, self.sn_k, self.sn2_q):
            sn.mark()
        u8 = isinstance(s, ops.Spikes)
        if not u8:
            s = s.view(wd * M * P, C)
        train_stats = bn_training(self.bn_q)
        qk_pre, part = ops.spike_linear(s, self._wqk(), stats=train_stats)         # [rows, 2C], one GEMM
        g = ops.qkgate(qk_pre.view(wd * M * P, 2 * C), _bn_of(self.bn_q), _bn_of(self.bn_k), self.positional_encoding,
                       self.sn_q.cfg(), wd, M, P, nH, partials=part, u8=u8)
        y, py = ops.spike_linear(g, self.proj.weight, self.proj.bias, stats=bn_training(self.proj_bn))
        return g, y, py

    def _core_generic(self, s, wd, M, P):
        """PSN / PLIF variant: the generic neuron kernels per site + library elementwise glue (fp32 spikes)."""
        C, nH = self.dim, self.num_heads
        if not isinstance(s, ops.Spikes):
            s = s.view(wd, M, P, C)
        self.sn_q.mark()
        q_pre, pq = ops.spike_linear(s, self.linear_q.weight, stats=bn_training(self.bn_q))
        k_pre, pk = ops.spike_linear(s, self.linear_k.weight, stats=bn_training(self.bn_k))
        q = ops.bn_neuron(q_pre.view(wd, M, P, C), _bn_of(self.bn_q), self.sn_q.cfg(), 0,
                          psn=self.sn_q.spiking_neuron if self.sn_q.is_psn else None, plif_w=self.sn_q.plif_w(), partials=pq)
        k = ops.bn_residual(k_pre.view(wd, M, P, C), _bn_of(self.bn_k),
                            self.positional_encoding.reshape(wd, 1, P, C).expand(wd, M, P, C), partials=pk)
        k = self.sn_k(k)
        att = self.sn2_q(q.reshape(wd, M, nH, -1, 32).sum(dim=-1, keepdim=True))
        g = k.reshape(M, nH, -1, 32) * att.reshape(M, nH, -1, 1)
        g = g.reshape(M, nH, wd, P, 32).permute(2, 0, 3, 1, 4).reshape(wd * M * P, C)
        return g, ops.spike_linear(g, self.proj.weight, self.proj.bias), None

    def forward(self, x, mask=None):
        """Reference call: x = x_windows (wd, B_, wh, ww, C) -> (x (B_, N, C), attention-score spikes)."""
        T, B_, P, C = self._geom_args(x)
        s = self.proj_sn(x.float().contiguous())
        g, y, py = self._core(s.view(T * B_ * P, C), T, B_, P)
        y = ops.bn_residual(y, _bn_of(self.proj_bn), partials=py)
        attn = self.attn_sn(g.view(T, B_, x.shape[2], x.shape[3], C))
        return y.view(B_, T * P, C), attn

    def fused(self, x, geom, alpha):
        """(B,D,H,W,C) -> (B,D,H,W,C): shortcut + DropPath(SSA(x)) with every index op folded in (:781-840)."""
        wd, C = geom.window[0], self.dim
        rows = geom.rows
        if not self.proj_sn.fusable:
            s = self.proj_sn(ops.window_gather(x, geom), u8=True)
        else:
            self.proj_sn.mark()
            s = ops.lif_window(x, geom, self.proj_sn.cfg(), u8=ops.spike_gemm_on())
        _, y, py = self._core(s, wd, geom.M, geom.P)
        return ops.window_scatter(y.view(rows, C), geom, res=x, bn_module=_bn_of(self.proj_bn), alpha=alpha, partials=py)


class Spiking_BN_WindowAttention3D(_WindowAttentionBase):
    """Q K^T V window attention with relative position bias, no softmax (reference :184-370):
    q,k,v = sn(bn(linear(x))); O = (scale*QK^T + bias + mask) @ V; proj -> proj_bn -> proj_sn."""
    sdsa = False

    def __init__(self, dim, window_size, pretrained_window_size, num_heads, version="swinv1", qkv_bias=False,
                 qk_scale=None, attn_drop=0.0, proj_drop=0.0, norm=None, **spiking_kwargs):
        super().__init__(dim, window_size, pretrained_window_size, num_heads, version, norm, spiking_kwargs)
        head_dim = dim // num_heads
        ws = self.window_size
        self.scale = 1 if spiking_kwargs["neuron_type"] in ["psn", "glif"] else (qk_scale or head_dim ** -0.5)
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * ws[0] - 1) * (2 * ws[1] - 1) * (2 * ws[2] - 1), num_heads))
        coords = torch.stack(torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), torch.arange(ws[2]),
                                            indexing="ij"))
        cf = torch.flatten(coords, 1)
        rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws[0] - 1
        rel[:, :, 1] += ws[1] - 1
        rel[:, :, 2] += ws[2] - 1
        rel[:, :, 0] *= (2 * ws[1] - 1) * (2 * ws[2] - 1)
        rel[:, :, 1] *= (2 * ws[2] - 1)
        self.register_buffer("relative_position_index", rel.sum(-1))
        for name in ("q", "k", "v"):
            setattr(self, f"linear_{name}", sj_layer.Linear(dim, dim, bias=False))
            setattr(self, f"bn_{name}", SpikingNormLayer(dim, ws[0], norm, spiking_kwargs["v_th"]))
            setattr(self, f"sn_{name}", Spiking_neuron(**spiking_kwargs))
        self.attn_drop = sj_layer.Dropout(attn_drop)
        self.attn_sn = Spiking_neuron(**spiking_kwargs)
        self.proj = sj_layer.Linear(dim, dim)
        self.proj_bn = SpikingNormLayer(dim, ws[0], norm, spiking_kwargs["v_th"])
        self.proj_sn = Spiking_neuron(**spiking_kwargs)
        self.proj_drop = sj_layer.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)

    def _attend(self, xin, wd, M, P, region, nW, want_attn=False):
        """xin [wd, M, P, C] -> proj output rows [wd*M*P, C] (q,k,v neurons + attention in one fused operator)."""
        ws = self.window_size
        if (wd, P) != (ws[0], ws[1] * ws[2]):
            raise NotImplementedError("relative position bias needs an unclamped window (stage >= window size)")
        sns = (self.sn_q, self.sn_k, self.sn_v)
        if self.sn_q.is_plif:
            raise NotImplementedError("Q K^T V window attention with ParametricLIF neurons is not built: the fused "
                                      "attention operator has no gradient path for the PLIF parameter")
        for sn in sns:
            sn.mark()
        pres = [ops.spike_linear(xin, getattr(self, f"linear_{n}").weight) for n in ("q", "k", "v")]
        psn = [sn.spiking_neuron for sn in sns] if self.sn_q.is_psn else None
        o, attn = ops.qktv_attention(pres[0], pres[1], pres[2], _bn_of(self.bn_q), _bn_of(self.bn_k), _bn_of(self.bn_v),
                                     self.relative_position_bias_table, self.sn_q.cfg(), region, wd, M, P, self.num_heads,
                                     nW, ws, float(self.scale), psn, want_attn)
        return ops.spike_linear(o, self.proj.weight, self.proj.bias, exact_input=False), attn   # O is real-valued

    def forward(self, x, mask=None, region=None, nW=1):
        """Reference call on x_windows (wd, B_, wh, ww, C).  The additive mask of the reference is
        expressed by region ids (ops.WindowGeom.region); passing a dense `mask` tensor is not supported."""
        if mask is not None:
            raise NotImplementedError("pass region ids (WindowGeom.region) instead of a dense attention mask")
        T, B_, P, C = self._geom_args(x)
        xin = x.contiguous().view(T, B_, P, C)
        if self.sdsa:
            xin = self.proj_sn(xin)
        y, attn = self._attend(xin, T, B_, P, region, nW, want_attn=True)
        if self.sdsa:
            y = ops.bn_residual(y, _bn_of(self.proj_bn))
        else:
            self.proj_sn.mark()
            y = ops.bn_neuron(y.view(T, B_, P, C), _bn_of(self.proj_bn), self.proj_sn.cfg(), 0,
                              psn=self.proj_sn.spiking_neuron if self.proj_sn.is_psn else None)
        return y.reshape(B_, T * P, C), attn

    def fused(self, x, geom, alpha):
        wd, C = geom.window[0], self.dim
        region = geom.region if geom.shifted else None
        if self.sdsa:
            if not self.proj_sn.fusable:
                xin = self.proj_sn(ops.window_gather(x, geom))
            else:
                self.proj_sn.mark()
                xin = ops.lif_window(x, geom, self.proj_sn.cfg())
        else:
            xin = ops.window_gather(x, geom)
        y, _ = self._attend(xin.view(wd, geom.M, geom.P, C), wd, geom.M, geom.P, region, geom.nW)
        if self.sdsa:
            return ops.window_scatter(y, geom, res=x, bn_module=_bn_of(self.proj_bn), alpha=alpha)
        self.proj_sn.mark()
        s = ops.bn_neuron(y.view(wd, geom.M, geom.P, C), _bn_of(self.proj_bn), self.proj_sn.cfg(), 0,
                          psn=self.proj_sn.spiking_neuron if self.proj_sn.is_psn else None)
        return ops.window_scatter(s.view(geom.rows, C), geom, res=x, bn_module=None, alpha=alpha)


class SDSA_WindowAttention3D(Spiking_BN_WindowAttention3D):
    """proj_sn first, no output neuron (reference :413-492)."""
    sdsa = True


# ---------------------------------------------------------------------------------------------
# block, merging, stage, backbone
# ---------------------------------------------------------------------------------------------
class Spiking_SwinTransformerBlock3D(nn.Module):
    """x = DropPath(SSA(x)) + x;  x = Mlp(x) + x  (reference :720-886)."""
    attn_module = Spiking_BN_WindowAttention3D
    mlp_module = Spiking_Mlp

    def __init__(self, dim, input_resolution, num_heads, window_size=(2, 7, 7), pretrained_window_size=(0, 0, 0),
                 shift_size=(0, 0, 0), mlp_ratio=4.0, version="swinv1", qkv_bias=True, qk_scale=None, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer="LN", use_checkpoint=False,
                 **spiking_kwargs):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        self.window_size, self.shift_size, self.mlp_ratio = tuple(window_size), tuple(shift_size), mlp_ratio
        self.use_checkpoint = use_checkpoint
        for s, w in zip(self.shift_size, self.window_size):
            assert 0 <= s < w, "shift_size must in 0-window_size"
        self.norm_layer = norm_layer
        if norm_layer in ["LN", "GN"]:
            raise NotImplementedError("LN/GN spike norm is not built on the B200 hot path (shipped configs use BN)")
        self.attn = self.attn_module(dim, window_size=self.window_size, pretrained_window_size=pretrained_window_size,
                                     num_heads=num_heads, version=version, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                     attn_drop=attn_drop, proj_drop=drop, norm=norm_layer, **dict(spiking_kwargs))
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.mlp = self.mlp_module(in_features=dim, hidden_features=int(dim * mlp_ratio), norm_layer=norm_layer,
                                   act_layer=act_layer, drop=drop, **spiking_kwargs)
        self.cnf = "ADD"

    def forward(self, x, mask_matrix=None, return_attention=False):
        """x: (B, D, H, W, C) contiguous.  mask_matrix is accepted for signature parity and ignored: the
        shift mask is derived from region ids inside the kernel."""
        if return_attention:
            raise NotImplementedError("return_attention: the reference's own log path mis-reads its axes "
                                      "(get_layer_attention_scores skips the rearrange, :1248-1264)")
        B, D, H, W, C = x.shape
        geom = ops.WindowGeom.get(B, D, H, W, self.window_size, self.shift_size, x.device)
        alpha = self.drop_path.scale(x) if isinstance(self.drop_path, DropPath) else None
        x = self.attn.fused(x.contiguous(), geom, alpha)
        return self.mlp.fused(x)

    def extra_repr(self):
        return (f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads}, "
                f"window_size={self.window_size}, shift_size={self.shift_size}, mlp_ratio={self.mlp_ratio}")


class MS_Spiking_SwinTransformerBlock3D(Spiking_SwinTransformerBlock3D):
    attn_module = Spiking_QK_WindowAttention3D
    mlp_module = MS_Spiking_Mlp


class SpikingPatchMerging(nn.Module):
    """2x2 gather -> Linear(4C -> 2C) -> BN -> neuron (reference :898-935)."""
    ms = False

    def __init__(self, input_resolution, dim, norm_layer="BN", **spiking_kwargs):
        super().__init__()
        self.input_resolution, self.dim = input_resolution, dim
        self.reduction = sj_layer.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = SpikingNormLayer(2 * dim, spiking_kwargs["num_steps"], norm_layer, spiking_kwargs["v_th"])
        self.sn = Spiking_neuron(**spiking_kwargs)

    def forward(self, x):
        """x: (B, D, H, W, C) -> (B, D, ceil(H/2), ceil(W/2), 2C)."""
        if self.ms:
            if not self.sn.fusable:
                s = self.sn(ops.lif_merge(x, self.sn.cfg(), apply_neuron=False), time_dim=1, u8=True)
            else:
                self.sn.mark()
                s = ops.lif_merge(x, self.sn.cfg(), u8=ops.spike_gemm_on())
            h, part = ops.spike_linear(s, self.reduction.weight, stats=bn_training(self.norm))
            return ops.bn_residual(h, _bn_of(self.norm), partials=part)
        g = ops.lif_merge(x, self.sn.cfg(), apply_neuron=False)
        self.sn.mark()
        return ops.bn_neuron(ops.spike_linear(g, self.reduction.weight), _bn_of(self.norm), self.sn.cfg(), 1,
                             psn=self.sn.spiking_neuron if self.sn.is_psn else None, plif_w=self.sn.plif_w())


class MS_SpikingPatchMerging(SpikingPatchMerging):
    """2x2 gather -> neuron -> Linear -> BN (reference :952-974)."""
    ms = True


class Spiking_Swin_BasicLayer(nn.Module):
    """One Swin stage (reference :995-1126)."""
    swin_block_type = Spiking_SwinTransformerBlock3D

    def __init__(self, dim, input_resolution, depth, num_heads, window_size=(1, 7, 7), pretrained_window_size=(1, 7, 7),
                 mlp_ratio=4.0, version="swinv1", qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 norm_layer="LN", downsample=None, use_checkpoint=False, **spiking_kwargs):
        super().__init__()
        self.dim, self.input_resolution, self.window_size = dim, input_resolution, tuple(window_size)
        self.shift_size = tuple(i // 2 for i in window_size)
        self.depth, self.use_checkpoint = depth, use_checkpoint
        self.swin_blocks = nn.ModuleList([
            self.swin_block_type(dim=dim, input_resolution=input_resolution, num_heads=num_heads,
                                 window_size=self.window_size, pretrained_window_size=pretrained_window_size,
                                 shift_size=(0, 0, 0) if (i % 2 == 0) else self.shift_size, mlp_ratio=mlp_ratio,
                                 version=version, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                                 drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                 norm_layer=norm_layer, use_checkpoint=use_checkpoint, **dict(spiking_kwargs))
            for i in range(depth)])
        self.downsample = downsample
        if self.downsample is not None:
            self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer, **dict(spiking_kwargs))

    def forward_cl(self, x):
        """Channels-last stage: x (B, D, H, W, C) -> (x_out (B, D, H', W', C'), x before merging)."""
        for blk in self.swin_blocks:
            x = blk(x)
        return (self.downsample(x) if self.downsample is not None else x), x

    def forward(self, x):
        """Reference signature: x (B, C, D, H, W) -> (x_out (B, C', D, H', W'), x (B, D, H, W, C))."""
        out, pre = self.forward_cl(x.permute(0, 2, 3, 4, 1).contiguous())
        return out.permute(0, 4, 1, 2, 3), pre

    def extra_repr(self):
        return f"dim={self.dim}, input_resolution={self.input_resolution}, depth={self.depth}"


class MS_Spiking_Swin_BasicLayer(Spiking_Swin_BasicLayer):
    swin_block_type = MS_Spiking_SwinTransformerBlock3D


class Spiking_SwinTransformer3D_v2(nn.Module):
    """Patch embedding + Swin stages (reference :1132-1284)."""
    swin_layer_type = Spiking_Swin_BasicLayer
    downsample_layer_type = SpikingPatchMerging

    def __init__(self, pretrained=None, pretrained2d=False, arc_type="swinv1", embed_type="PatchEmbedLocal",
                 img_size=(320, 480), patch_size=(4, 4, 4), in_chans=3, embed_dim=96, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=(2, 7, 7), pretrained_window_size=(2, 7, 7), mlp_ratio=4.0,
                 qkv_bias=True, qk_scale=0.125, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer="BN",
                 patch_norm=False, out_indices=(0, 1, 2, 3), frozen_stages=-1, use_checkpoint=False, norm=None,
                 **spiking_kwargs):
        super().__init__()
        self.pretrained, self.pretrained2d = pretrained, pretrained2d
        self.num_layers, self.embed_dim = len(depths), embed_dim
        self.patch_norm, self.frozen_stages = patch_norm, frozen_stages
        self.window_size, self.patch_size = window_size, patch_size
        self.out_indices, self.norm_layer = out_indices, norm_layer
        embed_cls = globals().get(embed_type)
        if embed_cls is None:
            raise NotImplementedError(f"patch embedding {embed_type!r} is not built; the shipped SNN configs use "
                                      "'MS_PED_Spiking_PatchEmbed_Conv_sfn'")
        self.patch_embed = embed_cls(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                     patch_norm=norm_layer if self.patch_norm else None, norm=norm, spiking_proj=True,
                                     **dict(spiking_kwargs))
        self.patches_resolution = self.patch_embed.patches_resolution
        self.pos_drop = sj_layer.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(self.swin_layer_type(
                dim=int(embed_dim * 2 ** i),
                input_resolution=(self.patches_resolution[0] // (2 ** i), self.patches_resolution[1] // (2 ** i)),
                depth=depths[i], num_heads=num_heads[i], window_size=tuple(window_size),
                pretrained_window_size=pretrained_window_size, mlp_ratio=mlp_ratio, version=arc_type, qkv_bias=qkv_bias,
                qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer=norm_layer,
                downsample=self.downsample_layer_type if i < self.num_layers - 1 else None,
                use_checkpoint=use_checkpoint, **dict(spiking_kwargs)))
        self.num_features = [int(embed_dim * 2 ** i) for i in range(self.num_layers)]
        if norm_layer in ["LN", "GN"]:
            raise NotImplementedError("LN/GN spike norm is not built on the B200 hot path")

    def forward_cl(self, x):
        """Channels-last stage loop: x (B, D, H, W, C) -> list of (B, D, Hi, Wi, Ci) per out index."""
        outs = []
        for i, lyr in enumerate(self.layers):
            x, pre = lyr.forward_cl(x)
            if i in self.out_indices:
                outs.append(pre)
        return outs

    def features_cl(self, x):
        """x: (B, bins, 2, H, W) voxels -> list of per-stage features (B, D, Hi, Wi, Ci), no layout copies."""
        if self.pos_drop.p != 0.0:
            raise NotImplementedError("pos_drop > 0 is not built (the reference passes drop_rate=0)")
        return self.forward_cl(self.patch_embed.forward_cl(x))

    def forward(self, x):
        """x: (B, bins, 2, H, W) voxels -> tuple of (B, Ci, D, Hi, Wi) views (reference :1223-1246)."""
        return tuple(o.permute(0, 4, 1, 2, 3) for o in self.features_cl(x))


class MS_Spiking_SwinTransformer3D_v2(Spiking_SwinTransformer3D_v2):
    """Spiking Swin Transformer 3D with MS shortcut (reference :1287-1292)."""
    swin_layer_type = MS_Spiking_Swin_BasicLayer
    downsample_layer_type = MS_SpikingPatchMerging
