"""Neuron switch, spike norm layer and the convolutional layers around the Swin encoder
(host-side mirror of reference models/STSwinNet_SNN/Spiking_modules.py).

Only what the three SNN FlowNets with the shipped patch embedding need is built:
Spiking_neuron (:26-99), SpikingNormLayer (:101-146), SpikingConvEncoderLayer (:250-296),
MS_SpikingConvEncoderLayer (:298-347), (MS_)SpikingTransposeDecoderLayer (:398-474),
(MS_)SpikingPredLayer (:568-647), SpikingPEDLayer (:772-825), SEWResBlock / MS_ResBlock
(:827-933), the residual feature generators (:935-973) and
MS_PED_Spiking_PatchEmbed_Conv_sfn (:1710-1790).  Convolutions stay cuDNN calls (out of the
hot-path scope); every neuron runs on the K1 kernel.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..sj import surrogate, neuron, functional, base, layer  # noqa: F401  (`surrogate` is eval()'d below)
from .. import capi, ops
from .Spiking_submodules import *  # noqa: F401,F403
from .Spiking_submodules import PSN, GatedLIFNode, SLTTLIFNode


class Spiking_neuron(nn.Module):
    """Switch over neuron types; `surrogate_fun` is a string eval()'d with `surrogate` in scope,
    as in the reference (:44).  State handling: the reference scripts call
    functional.reset_net(model) before every sample, so by default the membrane potential is
    not written back after a forward (saves 4 B/neuron); calling forward twice without a reset
    raises instead of silently diverging.  Set ``persist_state = True`` for spikingjelly's
    carry-over semantics."""
    persist_state = False

    def __init__(self, num_steps, spike_norm=None, neuron_type="plif", v_th=1.0, v_reset=0,
                 surrogate_fun="surrogate.ATan()", tau=2.0, detach_reset=True):
        super().__init__()
        assert neuron_type in ["lif", "if", "plif", "SLTTlif", "glif", "psn"]
        sf = eval(surrogate_fun) if isinstance(surrogate_fun, str) else surrogate_fun
        self.neuron_type = neuron_type
        if neuron_type == "lif":
            self.spiking_neuron = neuron.LIFNode(v_threshold=v_th, v_reset=v_reset, surrogate_function=sf, tau=tau,
                                                 detach_reset=detach_reset)
        elif neuron_type == "if":
            self.spiking_neuron = neuron.IFNode(v_threshold=v_th, v_reset=v_reset, surrogate_function=sf,
                                                detach_reset=detach_reset)
        elif neuron_type == "plif":
            self.spiking_neuron = neuron.ParametricLIFNode(v_threshold=v_th, v_reset=v_reset, surrogate_function=sf,
                                                           init_tau=tau, detach_reset=detach_reset)
        elif neuron_type == "psn":
            self.spiking_neuron = PSN(T=num_steps, surrogate_function=sf)
        elif neuron_type == "glif":
            self.spiking_neuron = GatedLIFNode(T=num_steps, surrogate_function=sf)
        else:
            self.spiking_neuron = SLTTLIFNode(v_threshold=v_th, v_reset=v_reset, surrogate_function=sf, tau=tau,
                                              detach_reset=detach_reset)
        self._dirty = False

    # ---- protocol -------------------------------------------------------------------------
    def reset(self):
        self._dirty = False

    @property
    def is_psn(self):
        return self.neuron_type == "psn"

    def cfg(self):
        return self.spiking_neuron.neuron_cfg()

    def plif_w(self):
        return getattr(self.spiking_neuron, "w", None)

    def mark(self):
        """Called by fused operators that run this neuron inside a kernel."""
        ops.tap_site(self.__dict__.get("_tap_name"))
        if self._dirty and not self.is_psn:
            raise RuntimeError("Spiking_neuron called twice without functional.reset_net(model); the reference "
                               "scripts reset before every sample (train_flow_parallel_supervised_SNN.py:238)")
        self._dirty = True

    @property
    def is_plif(self):
        return self.neuron_type == "plif"

    @property
    def fusable(self):
        """Can run inside the window / merge / QK-gate kernels (which take tau by value and have no gradient path for
        the PLIF parameter w)."""
        return self.neuron_type in ("lif", "if")

    def forward(self, x, time_dim=0, u8=False):
        """u8=True: return ops.Spikes (1 byte per spike) for a consumer that runs on the tcgen05 spike GEMM."""
        u8 = u8 and ops.spike_gemm_on()
        if self.is_psn:
            ops.tap_site(self.__dict__.get("_tap_name"))
            return ops.psn(x, self.spiking_neuron.weight, self.spiking_neuron.bias, self.cfg(), time_dim, u8=u8)
        if self.persist_state and time_dim == 0:
            if u8:
                raise RuntimeError("persist_state neurons return fp32 spikes")
            return self.spiking_neuron(x)
        self.mark()
        return ops.neuron(x, self.cfg(), time_dim, self.plif_w(), u8=u8)


class SpikingNormLayer(nn.Module):
    """Multi-step spike normalisation (reference :101-146).  The Swin blocks read `.norm_layer`'s
    tensors and fold the apply into the consuming kernel; called directly it normalises
    [T,B,C,H,W] like spikingjelly (statistics over T*B*H*W)."""

    def __init__(self, out_channels, num_steps, norm="BN", v_th=1.0):
        super().__init__()
        self.num_steps, self.norm = num_steps, norm
        if norm == "BN":
            self.norm_layer = layer.BatchNorm2d(out_channels)
        if norm == "BN_notrack":
            self.norm_layer = layer.BatchNorm2d(out_channels, track_running_stats=False)
        if norm == "GN":
            self.norm_layer = layer.GroupNorm(out_channels // 16, out_channels)
        if norm == "IN":
            self.norm_layer = layer.GroupNorm(out_channels, out_channels)
        if norm == "LN":
            self.norm_layer = layer.GroupNorm(1, out_channels)
        elif norm == "BNTT":
            self.norm_layer = nn.ModuleList([nn.BatchNorm2d(out_channels, eps=1e-4, momentum=0.1, affine=True)
                                             for _ in range(num_steps)])
        elif norm == "TDBN":
            self.norm_layer = layer.ThresholdDependentBatchNorm2d(alpha=1, v_th=v_th, num_features=out_channels)

    @property
    def is_batchnorm(self):
        return isinstance(self.norm_layer, nn.BatchNorm2d)

    def forward(self, x):
        if self.norm == "BNTT":
            return torch.cat([self.norm_layer[i](x[i]).unsqueeze(0) for i in range(self.num_steps)], dim=0)
        return self.norm_layer(x)



# ---------------------------------------------------------------------------------------------
# channels-last execution of the convolutional layers
# ---------------------------------------------------------------------------------------------
# Internally every layer below runs on ONE layout, (B, T, H, W, C) contiguous ("cl"): cuDNN gets NHWC
# tensors (tensor-core kernels without NCHW<->NHWC transforms), BatchNorm + neuron sites become
# channels-last rows for the fused K1/K2/K6 kernels (time axis = dim 1, no permute copies), and the
# Swin stages consume / produce the same layout.  `forward` keeps the reference's (T, B, C, H, W)
# signature as a thin adapter around `forward_cl`.
def to_cl(x):
    """(T, B, C, H, W) -> (B, T, H, W, C) contiguous."""
    return x.permute(1, 0, 3, 4, 2).contiguous()


def from_cl(y):
    """(B, T, H, W, C) -> logical (T, B, C, H, W) view."""
    return y.permute(1, 0, 4, 2, 3)


def u8_ok(conv, transposed=False):
    """Does this convolution run on the tcgen05 spike GEMM engine when fed 1-byte spikes?"""
    if not (ops.spike_gemm_on() and not transposed and isinstance(conv, nn.Conv2d) and conv.groups == 1
            and tuple(conv.dilation) == (1, 1)):
        return False
    if (tuple(conv.kernel_size), tuple(conv.stride), tuple(conv.padding)) == ((1, 1), (1, 1), (0, 0)) and conv.in_channels % 16 == 0:
        return True          # 1x1 = Linear on the rows; any Cout (ops.spike_linear pads the 2-channel heads)
    return ops.spike_conv_supported(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding)


def bn_training(norm_layer):
    """Does the BatchNorm inside this SpikingNormLayer / nn.BatchNorm2d use batch statistics now?"""
    bn = getattr(norm_layer, "norm_layer", norm_layer)
    return isinstance(bn, nn.modules.batchnorm._BatchNorm) and (bn.training or bn.running_mean is None)


def conv_cl(x, conv, spike_input, transposed=False, stats=None, weight=None):
    """x (B, T, H, W, Cin) fp32 or ops.Spikes -> (B, T, H', W', Cout); `conv` is an nn.Conv2d / nn.ConvTranspose2d parameter
    holder.  stats (bool, optional): when given, returns (y, BN partial sums of y or None) as ops.spike_linear does.
    weight: use this tensor instead of conv.weight (library path only; see padded_in_weight)."""
    if isinstance(x, ops.Spikes):
        sh = conv.stride[0]
        if tuple(conv.kernel_size) == (1, 1) and sh == 1:
            return ops.spike_linear(x, conv.weight.view(conv.weight.shape[0], -1), conv.bias, stats=stats)
        return ops.spike_conv_gemm(x, conv.weight, conv.bias, sh, conv.padding[0], stats=stats)
    y = _conv_cl_lib(x, conv, spike_input, transposed, weight)
    return y if stats is None else (y, None)


def padded_in_weight(conv, cin, transposed):
    """conv.weight with zero input-channel slices appended up to `cin` (the decoder input is concatenated with zero channels
    up to a multiple of 4 so that cuDNN's NHWC kernels take it without their own padding copy); autograd slices the gradient."""
    w = conv.weight
    have = w.shape[0] if transposed else w.shape[1]
    if cin == have:
        return None
    pad = (0, 0, 0, 0, 0, 0, 0, cin - have) if transposed else (0, 0, 0, 0, 0, cin - have)
    return F.pad(w, pad)


def _conv_cl_lib(x, conv, spike_input, transposed=False, weight=None):
    B, T, H, W, C = x.shape
    if weight is not None:
        x4 = x.view(B * T, H, W, C).permute(0, 3, 1, 2)
        y4 = ops.spike_conv2d(x4, weight, conv.bias, conv.stride, conv.padding, transposed,
                              conv.output_padding if transposed else 0, exact_input=spike_input)
        y = y4.permute(0, 2, 3, 1).contiguous()
        return y.view(B, T, y.shape[1], y.shape[2], y.shape[3])
    if (not transposed and not spike_input and C <= 4 and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (1, 1)
            and tuple(conv.padding) == (1, 1) and conv.weight.shape[0] % 4 == 0 and ops.GEMM_MODE != "fp32"):
        y = ops.conv3x3_small_cin(x.view(B * T, H, W, C), conv.weight, conv.bias)     # patch-embed head: direct kernel
        return y.view(B, T, H, W, y.shape[-1])
    if (not transposed and tuple(conv.kernel_size) == (1, 1) and tuple(conv.padding) == (0, 0) and conv.groups == 1):
        # 1x1 (strided) convolution == Linear on the (sub-sampled) channels-last rows: cuBLAS instead of cuDNN's
        # slow SIMT path (PED shortcut conv_res, 1x1 prediction heads)
        sh, sw = conv.stride
        xs = x if (sh, sw) == (1, 1) else x[:, :, ::sh, ::sw, :]
        return ops.spike_linear(xs, conv.weight.view(conv.weight.shape[0], C), conv.bias, exact_input=spike_input)
    x4 = x.view(B * T, H, W, C).permute(0, 3, 1, 2)                # logical NCHW with channels_last strides
    y4 = ops.spike_conv2d(x4, conv.weight, conv.bias, conv.stride, conv.padding, transposed,
                          conv.output_padding if transposed else 0, exact_input=spike_input)
    y = y4.permute(0, 2, 3, 1).contiguous()                        # no-op when cuDNN returned NHWC
    return y.view(B, T, y.shape[1], y.shape[2], y.shape[3])


def _bn_sn(h, norm_layer, sn, partials=None, u8=False):
    """neuron(BN(h)) on channels-last rows, time axis = dim 1."""
    sn.mark()
    return ops.bn_neuron(h, norm_layer.norm_layer, sn.cfg(), 1, psn=sn.spiking_neuron if sn.is_psn else None,
                         plif_w=sn.plif_w(), partials=partials, u8=u8 and ops.spike_gemm_on())


def _conv(cin, cout, k, stride, padding, bias):
    return nn.Sequential(layer.Conv2d(in_channels=cin, out_channels=cout, kernel_size=k, stride=stride, padding=padding,
                                      bias=bias))


class SpikingConvEncoderLayer(nn.Module):
    """conv -> norm -> neuron (SEW ordering, reference :250-296)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, spike_norm=None,
                 **spiking_kwargs):
        super().__init__()
        self.norm = spike_norm
        self.conv = _conv(in_channels, out_channels, kernel_size, stride, padding, self.norm is None)
        if self.norm is not None:
            self.norm_layer = SpikingNormLayer(out_channels, spiking_kwargs["num_steps"], self.norm,
                                               v_th=spiking_kwargs["v_th"])
        self.sn = Spiking_neuron(**spiking_kwargs)

    def forward_cl(self, x, spike_input=False, u8_out=False):
        """u8_out: emit the output spikes as ops.Spikes for a consumer on the tcgen05 spike GEMM."""
        if self.norm is not None and self.norm_layer.is_batchnorm:
            h, part = conv_cl(x, self.conv[0], spike_input, stats=bn_training(self.norm_layer))
            return _bn_sn(h, self.norm_layer, self.sn, part, u8_out)
        h = conv_cl(x, self.conv[0], spike_input)
        if self.norm is not None:
            h = to_cl(self.norm_layer(from_cl(h)))
        return self.sn(h, 1, u8=u8_out)

    def forward(self, x):
        return from_cl(self.forward_cl(to_cl(x)))


class MS_SpikingConvEncoderLayer(nn.Module):
    """[neuron ->] conv -> norm (membrane-shortcut ordering, reference :298-347)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, first_layer=True,
                 spike_norm=None, **spiking_kwargs):
        super().__init__()
        self.first_layer, self.norm = first_layer, spike_norm
        if not first_layer:
            self.sn = Spiking_neuron(**spiking_kwargs)
        self.conv = _conv(in_channels, out_channels, kernel_size, stride, padding, self.norm is None)
        self.conv[0].spike_input = not first_layer
        if self.norm is not None:
            self.norm_layer = SpikingNormLayer(out_channels, spiking_kwargs["num_steps"], self.norm,
                                               v_th=spiking_kwargs["v_th"])

    def forward_cl(self, x, spike_input=None):
        """spike_input: whether x is a spike tensor (exact in TF32).  Defaults to "a neuron runs first"; the patch
        embedding passes True for its first_layer=True conv, whose input is the head layer's spikes."""
        if not self.first_layer:
            x = self.sn(x, 1, u8=u8_ok(self.conv[0]))
        si = (not self.first_layer) if spike_input is None else spike_input
        if self.norm is None:
            return conv_cl(x, self.conv[0], si)
        if self.norm_layer.is_batchnorm:
            h, part = conv_cl(x, self.conv[0], si, stats=bn_training(self.norm_layer))
            return ops.bn_residual(h, self.norm_layer.norm_layer, partials=part)
        return to_cl(self.norm_layer(from_cl(conv_cl(x, self.conv[0], si))))

    def forward(self, x):
        return from_cl(self.forward_cl(to_cl(x)))


class SpikingTransposeDecoderLayer(nn.Module):
    """x2 (or x4) transposed-conv upsampling: deconv -> norm -> neuron (reference :398-459)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, spike_norm=None, scale=2, **spiking_kwargs):
        super().__init__()
        self.scale, self.norm = scale, spike_norm
        bias = self.norm is None
        if scale == 2:
            dc = layer.ConvTranspose2d(in_channels, out_channels, kernel_size, stride=2, padding=kernel_size // 2,
                                       output_padding=1, bias=bias)
        elif scale == 4:
            dc = layer.ConvTranspose2d(in_channels, out_channels, 7, stride=4, padding=2, output_padding=1, bias=bias)
        else:
            raise ValueError(scale)
        self.deconv = nn.Sequential(dc)
        if self.norm is not None:
            self.norm_layer = SpikingNormLayer(out_channels, spiking_kwargs["num_steps"], self.norm,
                                               v_th=spiking_kwargs["v_th"])
        self.sn = Spiking_neuron(**spiking_kwargs)

    def forward_cl(self, x, spike_input=True):
        h = conv_cl(x, self.deconv[0], spike_input, transposed=True, weight=padded_in_weight(self.deconv[0], x.shape[-1], True))
        if self.norm is not None:
            return _bn_sn(h, self.norm_layer, self.sn)
        return self.sn(h, 1)

    def forward(self, x):
        return from_cl(self.forward_cl(to_cl(x)))


class MS_SpikingTransposeDecoderLayer(SpikingTransposeDecoderLayer):
    """neuron -> deconv -> norm (reference :461-474)."""

    def forward_cl(self, x):
        dc = self.deconv[0]
        if ops.spike_gemm_on() and ops.spike_deconv_supported(dc, x.shape[-1]):
            # own engine: 1-byte spikes -> four parity-class implicit GEMMs, BN sums from their epilogues
            s = self.sn(x, 1, u8=True)
            if self.norm is None:
                return ops.spike_deconv(s, dc.weight, dc.bias)
            h, part = ops.spike_deconv(s, dc.weight, dc.bias, stats=bn_training(self.norm_layer))
            return ops.bn_residual(h, self.norm_layer.norm_layer, partials=part)
        h = conv_cl(self.sn(x, 1), dc, True, transposed=True, weight=padded_in_weight(dc, x.shape[-1], True))
        return ops.bn_residual(h, self.norm_layer.norm_layer) if self.norm is not None else h


class SpikingPredLayer(nn.Module):
    """1x1 prediction conv with bias (reference :568-605)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, **spiking_kwargs):
        super().__init__()
        self.norm = None
        self.conv = _conv(in_channels, out_channels, kernel_size, stride, kernel_size // 2, True)

    def forward_cl(self, x, spike_input=True):
        return conv_cl(x, self.conv[0], spike_input)

    def forward(self, x):
        return from_cl(self.forward_cl(to_cl(x)))


class MS_SpikingPredLayer(nn.Module):
    """neuron -> 1x1 prediction conv with bias (reference :607-647)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, **spiking_kwargs):
        super().__init__()
        self.norm = None
        self.sn = Spiking_neuron(**spiking_kwargs)
        self.conv = _conv(in_channels, out_channels, kernel_size, stride, kernel_size // 2, True)

    def forward_cl(self, x):
        return conv_cl(self.sn(x, 1, u8=u8_ok(self.conv[0])), self.conv[0], True)

    def forward(self, x):
        return from_cl(self.forward_cl(to_cl(x)))


class SpikingPEDLayer(nn.Module):
    """Patch embedding with deformed shortcut, /2 (reference :772-825): 1x1 s2 shortcut on the membrane
    input + neuron -> 3x3 s2 conv -> BN."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, norm=None,
                 patch_resolution=(120, 160), **spiking_kwargs):
        super().__init__()
        self.norm, self.patch = norm, patch_resolution
        bias = norm is None
        self.conv_res = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=2, padding=0, bias=bias)
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=1, bias=bias)
        if self.norm is not None:
            self.norm_layer = nn.BatchNorm2d(out_channels)
        self.sn = Spiking_neuron(**spiking_kwargs)

    def forward_cl(self, x):
        x_res = conv_cl(x, self.conv_res, False)                 # 1x1 stride-2 shortcut on the membrane input
        s = self.sn(x, 1, u8=u8_ok(self.conv))
        if self.norm is not None:
            y, part = conv_cl(s, self.conv, True, stats=bn_training(self.norm_layer))
            return ops.bn_residual(y, self.norm_layer, x_res, partials=part)
        return conv_cl(s, self.conv, True) + x_res

    def forward(self, x):
        return from_cl(self.forward_cl(to_cl(x))).contiguous()


def _connect(out, identity, fn):
    if fn == "ADD":
        return out + identity
    if fn in ("MUL", "AND"):
        return out * identity
    if fn == "OR":
        return surrogate.ATan(spiking=True)(out + identity)
    if fn == "NMUL":
        return identity * (1.0 - out)
    raise NotImplementedError(fn)


class _ResBlockBase(nn.Module):
    def __init__(self, in_channels, out_channels, stride=1, connect_function="ADD", spike_norm=None,
                 **spiking_kwargs):
        super().__init__()
        self.norm = spike_norm
        bias = self.norm is None
        self.conv1 = _conv(in_channels, out_channels, 3, stride, 1, bias)
        self.conv2 = _conv(in_channels, in_channels, 3, 1, 1, bias)
        if self.norm is not None:
            # reference quirk kept: SpikingNormLayer(out_channels, self.norm, v_th=...) => num_steps = "BN",
            # norm = default 'BN' (:848-849, :901-902)
            self.norm1 = SpikingNormLayer(out_channels, self.norm, v_th=spiking_kwargs["v_th"])
            self.norm2 = SpikingNormLayer(out_channels, self.norm, v_th=spiking_kwargs["v_th"])
        self.sn1 = Spiking_neuron(**spiking_kwargs)
        self.sn2 = Spiking_neuron(**spiking_kwargs)
        self.connect_function = connect_function


class SEWResBlock(_ResBlockBase):
    """conv-norm-neuron x2, spike-element-wise shortcut (reference :827-878)."""

    def forward_cl(self, x):
        ok2 = u8_ok(self.conv2[0])
        if self.norm is None:
            s = self.sn1(conv_cl(x, self.conv1[0], True), 1, u8=ok2)
            return _connect(self.sn2(conv_cl(s, self.conv2[0], True), 1), x, self.connect_function)
        h = conv_cl(x, self.conv1[0], True)                      # spikes (+ integer SEW sums): exact in TF32
        s = _bn_sn(h, self.norm1, self.sn1, u8=ok2)
        h, part = conv_cl(s, self.conv2[0], True, stats=bn_training(self.norm2))
        s = _bn_sn(h, self.norm2, self.sn2, part)
        return _connect(s, x, self.connect_function)

    def forward(self, x):
        return from_cl(self.forward_cl(to_cl(x)))


class MS_ResBlock(_ResBlockBase):
    """neuron-conv-norm x2, membrane shortcut (reference :880-933)."""

    def forward_cl(self, x):
        ok1, ok2 = u8_ok(self.conv1[0]), u8_ok(self.conv2[0])
        if self.norm is None:
            h = conv_cl(self.sn1(x, 1, u8=ok1), self.conv1[0], True)
            return _connect(conv_cl(self.sn2(h, 1, u8=ok2), self.conv2[0], True), x, self.connect_function)
        h, part = conv_cl(self.sn1(x, 1, u8=ok1), self.conv1[0], True, stats=bn_training(self.norm1))
        s = _bn_sn(h, self.norm1, self.sn2, part, ok2)
        h, part = conv_cl(s, self.conv2[0], True, stats=bn_training(self.norm2))
        if self.connect_function == "ADD":
            return ops.bn_residual(h, self.norm2.norm_layer, x, partials=part)
        h = ops.bn_residual(h, self.norm2.norm_layer, partials=part)
        return _connect(h, x, self.connect_function)

    def forward(self, x):
        return from_cl(self.forward_cl(to_cl(x)))


class spiking_residual_feature_generator(nn.Module):
    res_block_type = SEWResBlock

    def __init__(self, dim, norm, num_resblocks=4, cnt_fun="ADD", **spiking_kwargs):
        super().__init__()
        self.dim, self.num_resblocks = dim, num_resblocks
        self.resblocks = nn.ModuleList([
            self.res_block_type(dim, dim, stride=1, spike_norm=norm, connect_function=cnt_fun, **spiking_kwargs)
            for _ in range(num_resblocks)])

    def forward_cl(self, x):
        for blk in self.resblocks:
            x = blk.forward_cl(x)
        return x

    def forward(self, x):
        return from_cl(self.forward_cl(to_cl(x)))


class MS_spiking_residual_feature_generator(spiking_residual_feature_generator):
    res_block_type = MS_ResBlock


def regroup_bins_to_steps(x, num_bins, num_steps):
    """(B, bins, 2, H, W) -> (steps, B, num_ch, H, W) with new[:, i, ..., t] = x[:, (i//2)*steps + t, i%2]
    (reference :1772-1786), as one permute instead of a zero-fill + per-channel copy loop."""
    if x.size(1) > num_bins:
        x = x[:, :num_bins]
    B, _, _, H, W = x.shape
    g = num_bins // num_steps
    return x.reshape(B, g, num_steps, 2, H, W).permute(2, 0, 1, 3, 4, 5).reshape(num_steps, B, g * 2, H, W)


def regroup_bins_to_steps_cl(x, num_bins, num_steps):
    """Same regroup straight into the channels-last layout: (B, bins, 2, H, W) -> (B, steps, H, W, num_ch)."""
    if x.size(1) > num_bins:
        x = x[:, :num_bins]
    B, _, _, H, W = x.shape
    g = num_bins // num_steps
    return x.reshape(B, g, num_steps, 2, H, W).permute(0, 2, 4, 5, 1, 3).reshape(B, num_steps, H, W, g * 2).contiguous()


class MS_PED_Spiking_PatchEmbed_Conv_sfn(nn.Module):
    """Spiking patch embedding with PED, membrane shortcut (reference :1710-1790)."""
    use_MS = True
    num_res = 2
    first_conv_k = 3

    def __init__(self, img_size=(240, 320), patch_size=(2, 4, 4), in_chans=10, embed_dim=96, patch_norm=None, norm=None,
                 spiking_proj=False, spike_norm=None, **spiking_kwargs):
        super().__init__()
        self.patch_size, self.image_size = patch_size, img_size
        self.patches_resolution = [img_size[0] // patch_size[2] // 2, img_size[1] // patch_size[3] // 2]
        self.embed_dim, self.patch_norm = embed_dim, patch_norm
        self.num_bins, self.num_steps = in_chans, spiking_kwargs["num_steps"]
        self.num_ch = in_chans * 2 // self.num_steps
        self.spike_norm = spike_norm
        self.head = SpikingConvEncoderLayer(self.num_ch, embed_dim // 2, kernel_size=3, stride=1, padding=1,
                                            spike_norm=spike_norm, **spiking_kwargs)
        self.conv = MS_SpikingConvEncoderLayer(embed_dim // 2, embed_dim, kernel_size=self.first_conv_k, stride=2,
                                               padding=self.first_conv_k // 2, spike_norm=spike_norm, **spiking_kwargs)
        self.residual_encoding = MS_spiking_residual_feature_generator(dim=embed_dim, norm=spike_norm,
                                                                       num_resblocks=self.num_res, cnt_fun="ADD",
                                                                       **spiking_kwargs)
        self.proj = SpikingPEDLayer(embed_dim, embed_dim, kernel_size=3, stride=patch_size[2:], padding=1,
                                    norm=spike_norm, patch_resolution=self.patches_resolution, **spiking_kwargs)

    def forward_cl(self, x):
        """voxels (B, bins, 2, H, W) -> (B, T, H/4, W/4, embed_dim)."""
        if x.is_cuda and x.dtype == torch.float32 and not x.requires_grad:
            if x.size(1) > self.num_bins:
                x = x[:, :self.num_bins]
            x = ops.regroup_voxels(x, self.num_steps)             # one kernel, no permute copy (csrc/voxel_input.cu)
        else:
            x = regroup_bins_to_steps_cl(x, self.num_bins, self.num_steps)
        # real-valued voxel input: plain fp32 conv; its spikes feed `conv` as 1-byte Spikes when that runs on the spike GEMM
        x = self.head.forward_cl(x, spike_input=False, u8_out=u8_ok(self.conv.conv[0]))
        x = self.conv.forward_cl(x, spike_input=True)             # input = head's spikes
        x = self.residual_encoding.forward_cl(x)
        return self.proj.forward_cl(x)

    def forward(self, x):
        return from_cl(self.forward_cl(x)).contiguous()

    def extra_repr(self):
        return f" num_steps={self.num_steps}, patches_resolution={self.patches_resolution}"
