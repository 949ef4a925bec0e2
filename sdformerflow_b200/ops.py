"""Host-side operators: thin autograd wrappers over the C-ABI of libsdf_b200.so.

Every function here takes CUDA fp32 tensors, allocates outputs/workspaces with torch (the
library never allocates) and launches on torch's current stream.  There is NO CPU path: a
non-CUDA tensor raises.  Shapes follow the reference's tensors; comments cite the reference
lines each operator replaces.
"""
from dataclasses import dataclass
import math
import os
import weakref

import torch

from . import capi, gemm

N_PARTIAL = 444  # == sdf_partial_blocks(): up to 3 CTAs per SM x 148 SMs


class Spikes:
    """A 1-byte spike tensor on its way from the kernel that fires it to the GEMMs that consume it (the operand format of
    the tcgen05 kind::i8 spike GEMM: 1 B per spike in HBM instead of the reference's fp32 {0,1}).

    Autograd cannot carry an integer tensor, so the pair (token, grad) stands in for it: ``token`` is an fp32 scalar that
    the producing autograd node returns and every consuming node takes as an input — it only fixes the execution order of
    the backward pass — and the consumers' backward deposits dL/d(spikes) in ``grad`` (summed over consumers), which the
    producer's backward picks up.  Double backward through a Spikes edge is not supported."""
    __slots__ = ("data", "token", "grad")

    def __init__(self):
        self.data, self.token, self.grad = None, None, None

    @property
    def shape(self):
        return self.data.shape

    def add_grad(self, g):
        self.grad = g if self.grad is None else self.grad + g

    def take_grad(self):
        # called by the producer's backward, after every consumer's: the token is not needed any more, and dropping it breaks
        # the reference cycle producer node -> ctx.holder -> token -> grad_fn (= the producer node), which otherwise keeps
        # the 1-byte spikes of every eager step alive until Python's cyclic collector happens to run
        self.token = None
        g, self.grad = self.grad, None
        if g is None:       # no consumer produced a gradient (e.g. frozen weights downstream): zero
            g = torch.zeros(self.data.shape, device=self.data.device, dtype=torch.float32)
        return g.contiguous()


# ---- membrane taps (tests / monitors) ----------------------------------------------------------------------------------
# With ``TAP`` set to a dict, every neuron site that runs records its membrane potential after charge (fp32, the h of
# spikingjelly's neuronal_charge; the spikes are exactly h - v_th >= 0) under the name its Spiking_neuron module announced
# through ``tap_site`` just before the operator ran.  Production code never sets it.
TAP = None
_tap_pending = []


def tap_site(name):
    if TAP is not None:
        _tap_pending.append(name)


def _tap_take(n, *hs):
    """Record n membranes (or drop the pending names when this operator has no membrane output)."""
    if TAP is None:
        return
    names = [_tap_pending.pop(0) if _tap_pending else None for _ in range(n)]
    for nm, h in zip(names, hs):
        if nm is not None and h is not None:
            TAP[nm] = h


def _tapping():
    return TAP is not None


_zero_tokens = {}


def _zero_token(device):
    z = _zero_tokens.get(device)
    if z is None:
        z = _zero_tokens[device] = torch.zeros((), device=device, dtype=torch.float32)
    return z


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sdformerflow_b200: operators run on CUDA tensors only (there is no CPU fallback)")
        if t is not None and t.dtype != torch.float32 and t.dtype not in (torch.int32, torch.uint8, torch.bfloat16):
            raise RuntimeError(f"sdformerflow_b200: unexpected dtype {t.dtype}")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


_SPIKE_TORCH = {capi.SDF_SPIKE_F32: torch.float32, capi.SDF_SPIKE_U8: torch.uint8, capi.SDF_SPIKE_BF16: torch.bfloat16}


@dataclass(frozen=True)
class NeuronCfg:
    """Device-independent description of a spikingjelly neuron."""
    kind: int = capi.SDF_NEURON_LIF
    v_th: float = 1.0
    v_reset: object = 0.0          # None = soft reset
    tau: float = 2.0               # LIF tau; PLIF: 1/sigmoid(w) (set per call)
    detach_reset: bool = False
    surrogate: int = capi.SDF_SG_ATAN
    sg_alpha: float = 2.0

    def c(self, tau=None):
        return dict(kind=self.kind, hard_reset=0 if self.v_reset is None else 1,
                    detach_reset=1 if self.detach_reset else 0, surrogate=self.surrogate,
                    v_th=float(self.v_th), v_reset=0.0 if self.v_reset is None else float(self.v_reset),
                    tau=float(self.tau if tau is None else tau), sg_alpha=float(self.sg_alpha))


def seq_layout(shape, time_dim=0):
    """sdf_seq_layout for a contiguous tensor whose time axis is dim 0 ([T, ...]) or dim 1
    ((B, T, ...): replaces the x.permute(1,0,2,3,4) views of Spiking_swin_transformer3D.py:845)."""
    numel = 1
    for s in shape:
        numel *= s
    if time_dim == 0:
        T = shape[0]
        n = numel // T
        return dict(T=T, n_neurons=n, inner=n, stride_b=0, stride_t=n)
    if time_dim == 1:
        B, T = shape[0], shape[1]
        inner = numel // (B * T)
        return dict(T=T, n_neurons=B * inner, inner=inner, stride_b=T * inner, stride_t=inner)
    raise ValueError("time_dim must be 0 or 1")


# ---------------------------------------------------------------------------------------------
# BatchNorm bookkeeping shared by the fused operators
# ---------------------------------------------------------------------------------------------
class BNParams:
    """Bundle of a BatchNorm2d's tensors + hyper-parameters handed to fused operators."""
    __slots__ = ("weight", "bias", "running_mean", "running_var", "num_batches_tracked", "momentum", "eps",
                 "training")

    def __init__(self, bn: torch.nn.modules.batchnorm._BatchNorm):
        self.weight, self.bias = bn.weight, bn.bias
        self.running_mean, self.running_var = bn.running_mean, bn.running_var
        self.num_batches_tracked = bn.num_batches_tracked
        self.momentum = 0.1 if bn.momentum is None else bn.momentum
        self.eps = bn.eps
        # torch: batch statistics when training or when no running stats are tracked
        self.training = bn.training or bn.running_mean is None


NBT_DEFER = None      # list while a model forward collects its BatchNorm counters (defer_nbt): ~80 one-element adds -> one launch


class defer_nbt:
    """with ops.defer_nbt(): ... — the num_batches_tracked += 1 of every BatchNorm call inside is applied on exit as ONE
    torch._foreach_add_ (reference semantics: nn.BatchNorm2d increments its counter per training call)."""

    def __enter__(self):
        global NBT_DEFER
        self.outer = NBT_DEFER
        if self.outer is None:
            NBT_DEFER = []
        return self

    def __exit__(self, *exc):
        global NBT_DEFER
        if self.outer is None:
            pending, NBT_DEFER = NBT_DEFER, None
            if pending:
                uniq, counts = {}, {}
                for t in pending:
                    uniq[id(t)] = t
                    counts[id(t)] = counts.get(id(t), 0) + 1
                once = [uniq[k] for k in uniq if counts[k] == 1]
                if once:
                    torch._foreach_add_(once, 1)
                for k, c in counts.items():
                    if c > 1:
                        uniq[k].add_(c)
        return False


def _bn_forward_affine(u2d, rows, C, ld, bn: BNParams, partials=None):
    """-> scale, shift, mean, rstd (all [C]).  Train: batch stats + running update like torch.
    partials: per-block (sum, sum of squares) [N_PARTIAL, 2, C] already produced by the epilogue of the GEMM that wrote
    u2d (ops.spike_linear(..., stats=True)); without them a statistics pass over u2d runs here."""
    dev = u2d.device
    scale = torch.empty(C, device=dev, dtype=torch.float32)
    shift = torch.empty_like(scale)
    mean = torch.empty_like(scale)
    rstd = torch.empty_like(scale)
    if not bn.training:
        partials = None
    if bn.training:
        if partials is None:
            partials = torch.empty((N_PARTIAL, 2, C), device=dev, dtype=torch.float32)
            capi.call("sdf_bn_stats", capi.struct("sdf_bn_stats_args", x=_ptr(u2d), rows=rows, C=C, ld=ld,
                                                  partials=_ptr(partials), n_partial_blocks=N_PARTIAL, stream=_stream()),
                      algo_bytes=4 * rows * C)
        else:
            # [k * N_PARTIAL, 2, C]: k = 1 for a GEMM / convolution, 4 for the parity classes of a transposed convolution
            assert partials.shape[1:] == (2, C) and partials.shape[0] % N_PARTIAL == 0 and partials.is_contiguous()
        if bn.num_batches_tracked is not None:
            if NBT_DEFER is not None:
                NBT_DEFER.append(bn.num_batches_tracked)      # one multi-tensor add at the end of the model's forward
            else:
                bn.num_batches_tracked.add_(1)
    with torch.no_grad():
        capi.call("sdf_bn_finalize", capi.struct(
            "sdf_bn_finalize_args", partials=_ptr(partials), n_partial_blocks=partials.shape[0] if bn.training else 0,
            count=rows, C=C, weight=_ptr(bn.weight), bias=_ptr(bn.bias),
            running_mean=_ptr(bn.running_mean), running_var=_ptr(bn.running_var),
            momentum=float(bn.momentum), eps=float(bn.eps), training=1 if bn.training else 0,
            scale=_ptr(scale), shift=_ptr(shift), mean=_ptr(mean), rstd=_ptr(rstd), stream=_stream()))
    return scale, shift, mean, rstd


def _bn_bwd_coef(partials, rows, C, weight, mean, rstd, training):
    """partials = per-block (sum dy, sum dy*u) -> (coef [3, C] of du = a*dy + b*u + c, grad_weight, grad_bias)."""
    dev = partials.device
    gw = torch.empty(C, device=dev, dtype=torch.float32)
    gb = torch.empty_like(gw)
    coef = torch.empty((3, C), device=dev, dtype=torch.float32)
    capi.call("sdf_bn_bwd_finalize", capi.struct(
        "sdf_bn_bwd_finalize_args", partials=_ptr(partials), n_partial_blocks=N_PARTIAL, count=rows, C=C,
        weight=_ptr(weight), mean=_ptr(mean), rstd=_ptr(rstd), grad_weight=_ptr(gw), grad_bias=_ptr(gb),
        coef=_ptr(coef), training=1 if training else 0, stream=_stream()))
    return coef, gw, gb


def _bn_backward(partials, dy, u2d, ld_u, rows, C, weight, mean, rstd, training, need_du=True, du_ld=None):
    """partials = per-block (sum dy, sum dy*u).  -> du [rows, C], grad_weight, grad_bias."""
    dev = dy.device
    coef, gw, gb = _bn_bwd_coef(partials, rows, C, weight, mean, rstd, training)
    du = None
    if need_du:
        du = torch.empty((rows, C), device=dev, dtype=torch.float32)
        capi.call("sdf_bn_bwd_apply", capi.struct(
            "sdf_bn_bwd_apply_args", dy=_ptr(dy), u=_ptr(u2d), ld_u=ld_u, du=_ptr(du), ld_du=C, coef=_ptr(coef),
            rows=rows, C=C, stream=_stream()), algo_bytes=12 * rows * C)
    return du, gw, gb


# ---------------------------------------------------------------------------------------------
# K1/K2: multi-step neuron (optionally fused with the preceding BatchNorm)
# ---------------------------------------------------------------------------------------------
def _lif_fwd_raw(u, lay, cfg_c, spike_dtype, scale=None, shift=None, C=0, hw=1, want_h=False, v_init=None,
                 want_v=False):
    spike = torch.empty(u.shape, device=u.device, dtype=_SPIKE_TORCH[spike_dtype])
    h = torch.empty_like(u) if want_h else None
    v_final = torch.empty(lay["n_neurons"], device=u.device, dtype=torch.float32) if want_v else None
    capi.call("sdf_lif_fwd", capi.struct(
        "sdf_lif_fwd_args", u=_ptr(u), spike=_ptr(spike), h_seq=_ptr(h), v_init=_ptr(v_init), v_final=_ptr(v_final),
        scale=_ptr(scale), shift=_ptr(shift), C=C, hw=hw, lay=lay, neuron=cfg_c, spike_dtype=spike_dtype,
        stream=_stream()), algo_bytes=u.numel() * (4 + spike.element_size()))
    return spike, h, v_final


class _NeuronFn(torch.autograd.Function):
    """Plain multi-step LIF/IF/PLIF: spikingjelly LIFNode.multi_step_forward via Spiking_neuron.forward
    (reference Spiking_modules.py:98-99)."""

    @staticmethod
    def forward(ctx, u, plif_w, cfg, time_dim, v_init, want_state, holder=None):
        _need_cuda(u)
        u = u.contiguous()
        lay = seq_layout(u.shape, time_dim)
        cc = cfg.c(plif_tau(plif_w) if cfg.kind == capi.SDF_NEURON_PLIF else None)
        dt = capi.SDF_SPIKE_F32 if holder is None else capi.SDF_SPIKE_U8
        spike, h_tap, v_final = _lif_fwd_raw(u, lay, cc, dt, v_init=v_init, want_v=want_state, want_h=_tapping())
        _tap_take(1, h_tap)
        ctx.save_for_backward(u, plif_w, v_init)
        ctx.lay, ctx.cc, ctx.cfg, ctx.holder = lay, cc, cfg, holder
        if holder is not None:
            holder.data = spike
            spike = u.new_empty(())          # the token (see Spikes)
        if want_state:
            ctx.mark_non_differentiable(v_final)
            return spike, v_final
        return spike, None

    @staticmethod
    def backward(ctx, gs, _gv):
        u, plif_w, v_init = ctx.saved_tensors
        gs = gs.contiguous() if ctx.holder is None else ctx.holder.take_grad()
        gu = torch.empty_like(u)
        plif_part = None
        if ctx.cfg.kind == capi.SDF_NEURON_PLIF:
            plif_part = torch.empty(N_PARTIAL, device=u.device, dtype=torch.float32)
        capi.call("sdf_lif_bwd", capi.struct(
            "sdf_lif_bwd_args", u=_ptr(u), grad_spike=_ptr(gs), grad_u=_ptr(gu), v_init=_ptr(v_init),
            plif_partials=_ptr(plif_part), n_partial_blocks=N_PARTIAL, C=0, hw=1, lay=ctx.lay, neuron=ctx.cc,
            stream=_stream()), algo_bytes=12 * u.numel())
        gw = None
        if plif_part is not None:
            s = torch.sigmoid(plif_w.detach())
            gw = (plif_part.sum() * s * (1 - s)).reshape(plif_w.shape)
        return gu, gw, None, None, None, None, None


_plif_tau_cache = {}


def plif_tau(plif_w):
    """1 / sigmoid(w) of a ParametricLIFNode as a host float (the kernels take tau by value).  One device->host read per
    parameter version, so PLIF models are not CUDA-graph capturable; lif / if / psn never come here."""
    key = (plif_w.data_ptr(), plif_w._version, gemm.weights_epoch())      # epoch: fused optimizers do not bump _version
    hit = _plif_tau_cache.get(id(plif_w))
    if hit is not None and hit[0] == key and hit[2]() is plif_w:       # the entry of THIS live tensor, not of a freed one
        return hit[1]
    tau = 1.0 / torch.sigmoid(plif_w.detach()).item()
    wid = id(plif_w)
    _plif_tau_cache[wid] = (key, tau, weakref.ref(plif_w, lambda _r, wid=wid: _plif_tau_cache.pop(wid, None)
                                                  if (wid in _plif_tau_cache and _plif_tau_cache[wid][2] is _r) else None))
    return tau


def neuron(u, cfg: NeuronCfg, time_dim=0, plif_w=None, v_init=None, want_state=False, u8=False):
    """spikes of a multi-step neuron over dim `time_dim` — fp32 {0,1}, or a Spikes (1 byte each) with u8=True;
    optionally the final membrane."""
    if v_init is not None and v_init.requires_grad:
        # the kernels treat the carried membrane as a constant; silently dropping its gradient would be a wrong answer.
        # The reference never needs it: functional.reset_net runs before every forward (train_…_SNN.py:247).
        raise RuntimeError("ops.neuron: gradient w.r.t. the initial membrane (state carried across calls) is not "
                           "implemented; detach() the state or reset the net between calls")
    holder = Spikes() if u8 else None
    spike, v = _NeuronFn.apply(u, plif_w, cfg, time_dim, v_init, want_state, holder)
    if u8:
        holder.token = spike
        spike = holder
    return (spike, v) if want_state else spike


def neuron_debug(u, cfg: NeuronCfg, time_dim=0, scale=None, shift=None, C=0, hw=1, spike_dtype=capi.SDF_SPIKE_F32):
    """(spikes, membrane-after-charge h) without autograd — for parity tests and monitors."""
    _need_cuda(u)
    u = u.contiguous()
    spike, h, _ = _lif_fwd_raw(u, seq_layout(u.shape, time_dim), cfg.c(), spike_dtype, scale, shift, C, hw, want_h=True)
    return spike, h


class _PSNFn(torch.autograd.Function):
    """Parallel spiking neuron: s = heaviside(W x + b) over the time axis
    (reference Spiking_submodules.py:207-211: addmm + surrogate)."""

    @staticmethod
    def forward(ctx, u, weight, bias, cfg, time_dim, holder=None):
        _need_cuda(u, weight, bias)
        u = u.contiguous()
        lay = seq_layout(u.shape, time_dim)
        _check_psn(weight, bias, lay)
        dt = capi.SDF_SPIKE_F32 if holder is None else capi.SDF_SPIKE_U8
        spike = torch.empty(u.shape, device=u.device, dtype=_SPIKE_TORCH[dt])
        w, b = weight.detach().contiguous(), bias.detach().contiguous()
        h_tap = torch.empty_like(u) if _tapping() else None
        capi.call("sdf_psn_fwd", capi.struct(
            "sdf_psn_fwd_args", u=_ptr(u), spike=_ptr(spike), h_seq=_ptr(h_tap), weight=_ptr(w), bias=_ptr(b), C=0, hw=1, lay=lay,
            spike_dtype=dt, stream=_stream()), algo_bytes=u.numel() * (4 + spike.element_size()))
        _tap_take(1, h_tap)
        ctx.save_for_backward(u, w, b)
        ctx.lay, ctx.cfg, ctx.holder = lay, cfg, holder
        if holder is not None:
            holder.data = spike
            return u.new_empty(())
        return spike

    @staticmethod
    def backward(ctx, gs):
        u, w, b = ctx.saved_tensors
        gs = gs.contiguous() if ctx.holder is None else ctx.holder.take_grad()
        T, n = ctx.lay["T"], ctx.lay["n_neurons"]
        gu = torch.empty_like(u)
        if _psn_fused_wgrad_ok(ctx.lay, u, gs, gu):
            pg = _psn_bwd_fused(T, u, dict(
                u=_ptr(u), grad_spike=_ptr(gs), grad_u=_ptr(gu), grad_h=None, x_out=None, weight=_ptr(w), bias=_ptr(b), C=0, hw=1,
                lay=ctx.lay, surrogate=ctx.cfg.surrogate, sg_alpha=float(ctx.cfg.sg_alpha), stream=_stream()))
            if pg is not None:
                return gu, pg[0], pg[1], None, None, None
        gh = torch.empty((T, n), device=u.device, dtype=torch.float32)
        xo = u.view(T, n) if ctx.lay["stride_b"] == 0 else torch.empty((T, n), device=u.device, dtype=torch.float32)
        capi.call("sdf_psn_bwd", capi.struct(
            "sdf_psn_bwd_args", u=_ptr(u), grad_spike=_ptr(gs), grad_u=_ptr(gu), grad_h=_ptr(gh),
            x_out=None if ctx.lay["stride_b"] == 0 else _ptr(xo), weight=_ptr(w), bias=_ptr(b), C=0, hw=1,
            lay=ctx.lay, surrogate=ctx.cfg.surrogate, sg_alpha=float(ctx.cfg.sg_alpha), stream=_stream()))
        g_w, g_b = _psn_param_grads(gh, xo.contiguous() if not xo.is_contiguous() else xo)
        return gu, g_w, g_b, None, None, None


N_PSN_WG = 2048      # rows of the PSN parameter-gradient partial buffer (>= blocks of any sdf_psn_bwd launch)


def _psn_bwd_fused(T, u, kw):
    """sdf_psn_bwd with the parameter gradients accumulated in the same pass -> (dW [T,T], db [T,1]), or None when the launch
    geometry of this layout cannot take it (the caller then runs the two-kernel path; nothing was launched)."""
    wpart = torch.empty((N_PSN_WG, T * T + T), device=u.device, dtype=torch.float32)
    try:
        capi.call("sdf_psn_bwd", capi.struct("sdf_psn_bwd_args", wgrad_partials=_ptr(wpart), n_wgrad_blocks=N_PSN_WG, **kw),
                  algo_bytes=12 * u.numel())
    except RuntimeError as e:
        if "wgrad_partials" not in str(e):
            raise
        return None
    tot = wpart.sum(0)
    return tot[:T * T].view(T, T), tot[T * T:].view(T, 1)


def _psn_fused_wgrad_ok(lay, *tensors):
    """Can sdf_psn_bwd accumulate dW / db itself (its vector path: T in {2,4,5,10}, every stride a multiple of 4 elements,
    16-byte aligned buffers)?  Otherwise grad_h / x go through HBM to sdf_psn_wgrad."""
    return (lay["T"] in (2, 4, 5, 10) and lay["n_neurons"] % 4 == 0 and lay["inner"] % 4 == 0 and lay["stride_b"] % 4 == 0
            and lay["stride_t"] % 4 == 0 and all(t.data_ptr() % 16 == 0 for t in tensors))


def _psn_param_grads(gh, xo):
    """(dW [T,T], db [T,1]) of a PSN from the kernel's grad_h and x: own strided-reduction kernel for the T the models use
    (a [T, n] x [n, T] library GEMM with n ~ 1e8 cost 55-72 ms per training step), library fallback otherwise."""
    T, n = gh.shape
    if T in (2, 4, 5, 10):
        part = torch.empty((N_PARTIAL, T * T + T), device=gh.device, dtype=torch.float32)
        capi.call("sdf_psn_wgrad", capi.struct("sdf_psn_wgrad_args", grad_h=_ptr(gh), x=_ptr(xo), partials=_ptr(part),
                                               n_partial_blocks=N_PARTIAL, T=T, n_neurons=n, stream=_stream()),
                  algo_bytes=8 * T * n)
        tot = part.sum(0)
        return tot[:T * T].view(T, T), tot[T * T:].view(T, 1)
    with _tf32(False):      # other T: rare (no shipped config), keep the library GEMM in true fp32
        return gh @ xo.t(), gh.sum(1, keepdim=True)


def _check_psn(weight, bias, lay):
    """The kernels index the [T, T] matrix with the layout's T: a PSN built for another number of steps (window depth
    clamped, num_steps != tensor's time extent) must fail like the reference's addmm shape error, not read out of bounds."""
    T = lay["T"]
    if tuple(weight.shape) != (T, T) or bias.numel() != T:
        raise RuntimeError(f"PSN: weight {tuple(weight.shape)} / bias {tuple(bias.shape)} do not match the input's T={T}")


def psn(u, weight, bias, cfg: NeuronCfg, time_dim=0, u8=False):
    if not u8:
        return _PSNFn.apply(u, weight, bias, cfg, time_dim, None)
    holder = Spikes()
    holder.token = _PSNFn.apply(u, weight, bias, cfg, time_dim, holder)
    return holder


class _BNNeuronFn(torch.autograd.Function):
    """y = neuron(BN(u)) on channels-last rows: `sn(bn(linear(x)).permute..)` sites, e.g. reference
    Spiking_swin_transformer3D.py:171-174 (bn1 -> sn2), :310-311, :933-934."""

    @staticmethod
    def forward(ctx, u, weight, bias, bn, cfg, time_dim, psn_w, psn_b, plif_w, partials, holder):
        _need_cuda(u)
        u = u.contiguous()
        C = u.shape[-1]
        rows = u.numel() // C
        scale, shift, mean, rstd = _bn_forward_affine(u, rows, C, C, bn, partials)
        lay = seq_layout(u.shape, time_dim)
        dt = capi.SDF_SPIKE_F32 if holder is None else capi.SDF_SPIKE_U8
        if psn_w is None:
            cc = cfg.c(plif_tau(plif_w) if cfg.kind == capi.SDF_NEURON_PLIF else None)
            spike, h_tap, _ = _lif_fwd_raw(u, lay, cc, dt, scale, shift, C, 1, want_h=_tapping())
        else:
            cc = None
            _check_psn(psn_w, psn_b, lay)
            spike = torch.empty(u.shape, device=u.device, dtype=_SPIKE_TORCH[dt])
            h_tap = torch.empty_like(u) if _tapping() else None
            capi.call("sdf_psn_fwd", capi.struct(
                "sdf_psn_fwd_args", u=_ptr(u), spike=_ptr(spike), h_seq=_ptr(h_tap), weight=_ptr(psn_w), bias=_ptr(psn_b),
                scale=_ptr(scale), shift=_ptr(shift), C=C, hw=1, lay=lay, spike_dtype=dt,
                stream=_stream()), algo_bytes=u.numel() * (4 + spike.element_size()))
        _tap_take(1, h_tap)
        ctx.save_for_backward(u, weight, scale, shift, mean, rstd, psn_w, psn_b, plif_w)
        ctx.lay, ctx.cc, ctx.cfg, ctx.training, ctx.rows, ctx.C = lay, cc, cfg, bn.training, rows, C
        ctx.holder = holder
        if holder is not None:
            holder.data = spike
            return u.new_empty(())
        return spike

    @staticmethod
    def backward(ctx, gs):
        u, weight, scale, shift, mean, rstd, psn_w, psn_b, plif_w = ctx.saved_tensors
        gs = gs.contiguous() if ctx.holder is None else ctx.holder.take_grad()
        rows, C = ctx.rows, ctx.C
        dev = u.device
        partials = torch.empty((N_PARTIAL, 2, C), device=dev, dtype=torch.float32)
        g_psn_w = g_psn_b = g_plif = None
        if psn_w is None:
            # (a two-phase variant — statistics-only pass, then a second walk of the recurrence applying the BatchNorm
            # backward in registers via sdf_lif_bwd_args.bn_coef, 20 B instead of 24 B per neuron-timestep — measured
            # SLOWER on B200: K2 is register/latency bound at T = 10, sdf_lif_bwd 4.5 -> 7.7 ms vs sdf_bn_bwd_apply
            # 4.6 -> 2.6 ms per step; kept in the C-ABI, not used here)
            plif_part = None
            if ctx.cfg.kind == capi.SDF_NEURON_PLIF:
                plif_part = torch.empty(N_PARTIAL, device=dev, dtype=torch.float32)
            dx = torch.empty_like(u)
            capi.call("sdf_lif_bwd", capi.struct(
                "sdf_lif_bwd_args", u=_ptr(u), grad_spike=_ptr(gs), grad_u=None, grad_x=_ptr(dx), scale=_ptr(scale),
                shift=_ptr(shift), bn_partials=_ptr(partials), plif_partials=_ptr(plif_part), n_partial_blocks=N_PARTIAL,
                C=C, hw=1, lay=ctx.lay, neuron=ctx.cc, stream=_stream()), algo_bytes=12 * u.numel())
            if plif_part is not None:
                sg = torch.sigmoid(plif_w.detach())
                g_plif = (plif_part.sum() * sg * (1 - sg)).reshape(plif_w.shape)
            du, gw, gb = _bn_backward(partials, dx, u, C, rows, C, weight, mean, rstd, ctx.training)
            return du.view(u.shape), gw, gb, None, None, None, g_psn_w, g_psn_b, g_plif, None, None
        dx = torch.empty_like(u)
        T, n = ctx.lay["T"], ctx.lay["n_neurons"]
        pg = None
        if _psn_fused_wgrad_ok(ctx.lay, u, gs, dx):
            pg = _psn_bwd_fused(T, u, dict(
                u=_ptr(u), grad_spike=_ptr(gs), grad_u=None, grad_x=_ptr(dx), grad_h=None, x_out=None, weight=_ptr(psn_w),
                bias=_ptr(psn_b), scale=_ptr(scale), shift=_ptr(shift), bn_partials=_ptr(partials), n_partial_blocks=N_PARTIAL,
                C=C, hw=1, lay=ctx.lay, surrogate=ctx.cfg.surrogate, sg_alpha=float(ctx.cfg.sg_alpha), stream=_stream()))
        if pg is not None:
            g_psn_w, g_psn_b = pg
        else:
            gh = torch.empty((T, n), device=dev, dtype=torch.float32)
            xo = torch.empty((T, n), device=dev, dtype=torch.float32)
            capi.call("sdf_psn_bwd", capi.struct(
                "sdf_psn_bwd_args", u=_ptr(u), grad_spike=_ptr(gs), grad_u=None, grad_x=_ptr(dx), grad_h=_ptr(gh),
                x_out=_ptr(xo), weight=_ptr(psn_w), bias=_ptr(psn_b), scale=_ptr(scale), shift=_ptr(shift),
                bn_partials=_ptr(partials), n_partial_blocks=N_PARTIAL, C=C, hw=1, lay=ctx.lay,
                surrogate=ctx.cfg.surrogate, sg_alpha=float(ctx.cfg.sg_alpha), stream=_stream()))
            g_psn_w, g_psn_b = _psn_param_grads(gh, xo)
        du, gw, gb = _bn_backward(partials, dx, u, C, rows, C, weight, mean, rstd, ctx.training)
        return du.view(u.shape), gw, gb, None, None, None, g_psn_w, g_psn_b, g_plif, None, None


def _seq_to_layout(seq, shape, lay):
    """[T, n] contiguous -> tensor laid out like the original input (inverse of the kernel's addressing)."""
    T = lay["T"]
    if lay["stride_b"] == 0:
        return seq.reshape(shape).contiguous()
    B = lay["n_neurons"] // lay["inner"]
    return seq.view(T, B, lay["inner"]).permute(1, 0, 2).contiguous().view(shape)


def bn_neuron(u, bn_module, cfg: NeuronCfg, time_dim=0, psn=None, plif_w=None, partials=None, u8=False):
    """neuron(BN(u)); u is channels-last [..., C]; BN statistics over all leading dims (the
    spikingjelly multi-step BN flattens (T,B): SURVEY.md Appendix A).  partials: BN partial sums from the GEMM that
    produced u (skips the statistics pass); u8: return a Spikes (1 byte per spike) for the spike GEMM."""
    bn = BNParams(bn_module)
    pw, pb = (psn.weight, psn.bias) if psn is not None else (None, None)
    if cfg.kind == capi.SDF_NEURON_PLIF and plif_w is None:
        raise RuntimeError("bn_neuron: a ParametricLIFNode needs its parameter w (plif_w)")
    holder = Spikes() if u8 else None
    out = _BNNeuronFn.apply(u, bn.weight, bn.bias, bn, cfg, time_dim, pw, pb, plif_w, partials, holder)
    if u8:
        holder.token = out
        return holder
    return out


class _BNResidualFn(torch.autograd.Function):
    """out = res + BN(u) on channels-last rows (MS MLP tail: Spiking_swin_transformer3D.py:176-178 + :845)."""

    @staticmethod
    def forward(ctx, u, res, weight, bias, bn, partials=None):
        _need_cuda(u, res)
        u = u.contiguous()
        C = u.shape[-1]
        rows = u.numel() // C
        scale, shift, mean, rstd = _bn_forward_affine(u, rows, C, C, bn, partials)
        out = torch.empty_like(u)
        if res is not None:
            res = res.contiguous()
        capi.call("sdf_bn_apply", capi.struct(
            "sdf_bn_apply_args", u=_ptr(u), ld_u=C, res=_ptr(res), out=_ptr(out), scale=_ptr(scale), shift=_ptr(shift),
            rows=rows, C=C, stream=_stream()), algo_bytes=(8 if res is None else 12) * rows * C)
        ctx.save_for_backward(u, weight, mean, rstd)
        ctx.training, ctx.rows, ctx.C, ctx.has_res = bn.training, rows, C, res is not None
        return out

    @staticmethod
    def backward(ctx, go):
        u, weight, mean, rstd = ctx.saved_tensors
        go = go.contiguous()
        rows, C = ctx.rows, ctx.C
        partials = torch.empty((N_PARTIAL, 2, C), device=u.device, dtype=torch.float32)
        capi.call("sdf_bn_bwd_reduce", capi.struct(
            "sdf_bn_bwd_reduce_args", dy=_ptr(go), u=_ptr(u), ld_u=C, rows=rows, C=C, partials=_ptr(partials),
            n_partial_blocks=N_PARTIAL, stream=_stream()), algo_bytes=8 * rows * C)
        du, gw, gb = _bn_backward(partials, go, u, C, rows, C, weight, mean, rstd, ctx.training)
        return du.view(u.shape), (go if ctx.has_res else None), gw, gb, None, None


def bn_residual(u, bn_module, res=None, partials=None):
    bn = BNParams(bn_module)
    return _BNResidualFn.apply(u, res, bn.weight, bn.bias, bn, partials)


# ---------------------------------------------------------------------------------------------
# fp32-faithful spike GEMM / convolution on tensor cores (library calls)
# ---------------------------------------------------------------------------------------------
# A spike operand is exactly representable in TF32 (so is a small-integer SEW residual sum), hence
# every product spike * w_tf32 is exact and accumulates in fp32.  Splitting W = hi + lo with both
# parts TF32-representable (|W - hi - lo| <= 2^-22 |W|) turns one fp32 GEMM/conv on the SIMT pipe
# into two TF32 tensor-core calls with fp32-grade results.  Backward runs single-pass TF32 (the
# reference trains under fp16 autocast, train_flow_parallel_supervised_SNN.py:248, so this is at
# least as precise).  GEMM_MODE = "fp32" restores plain SIMT fp32 everywhere.
GEMM_MODE = "tf32x2"
# True: spike tensors between kernels are 1-byte Spikes and Linear / Conv on them run on the library's own tcgen05 + TMA
# GEMM engine (csrc/spike_gemm.cu, csrc/spike_wgrad.cu); False: fp32 spikes through the cuBLAS / cuDNN TF32 x 2 path above.
USE_SPIKE_GEMM = True


def spike_gemm_on():
    return USE_SPIKE_GEMM and GEMM_MODE != "fp32"


def split_tf32(w):
    """w (fp32) -> hi, lo, both exactly representable in TF32, hi + lo == w up to 2^-22 |w| (one kernel)."""
    w = w.contiguous()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    capi.call("sdf_split_tf32", capi.struct("sdf_split_tf32_args", w=_ptr(w), hi=_ptr(hi), lo=_ptr(lo), n=w.numel(),
                                            stream=_stream()))
    return hi, lo


class _tf32:
    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = self.on
        torch.backends.cudnn.allow_tf32 = self.on

    def __exit__(self, *a):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.prev


class _SpikeLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, weight, bias):
        s2 = s.reshape(-1, s.shape[-1])
        hi, lo = split_tf32(weight.detach())
        with _tf32(True):
            y = torch.addmm(bias.detach(), s2, hi.t()) if bias is not None else torch.mm(s2, hi.t())
            y.addmm_(s2, lo.t())
        ctx.save_for_backward(s, weight)
        ctx.has_bias = bias is not None
        return y.view(*s.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, g):
        s, weight = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        s2 = s.reshape(-1, s.shape[-1])
        gs = gw = gb = None
        with _tf32(True):
            if ctx.needs_input_grad[0]:
                gs = torch.mm(g2, weight).view(s.shape)
            if ctx.needs_input_grad[1]:
                gw = torch.mm(g2.t(), s2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.sum(0)
        return gs, gw, gb


class _PlainLinearFn(torch.autograd.Function):
    """Real-valued operand: true fp32 forward (SIMT), single-pass TF32 backward."""

    @staticmethod
    def forward(ctx, s, weight, bias):
        with _tf32(False):
            y = torch.nn.functional.linear(s, weight, bias)
        ctx.save_for_backward(s, weight)
        ctx.has_bias = bias is not None
        return y

    backward = _SpikeLinearFn.backward


class _SpikeGemmFn(torch.autograd.Function):
    """y = spikes @ W^T (+ b) on the tcgen05 engine: kind::i8 forward on the 1-byte spikes (gemm.spike_gemm_fwd, exact
    integer accumulation of three weight digit planes, BN partial sums from the epilogue), TF32 data gradient
    (gemm.gemm_tf32) and MN-major bf16 hi/lo weight gradient (gemm.spike_wgrad).  Reference: sj_layer.Linear on spike tensors,
    Spiking_swin_transformer3D.py:126-131,267-290,632-652,909."""

    @staticmethod
    def forward(ctx, token, weight, bias, holder, want_stats):
        ctx.set_materialize_grads(False)      # the BN partial sums get no gradient: do not zero-fill one per call
        a = holder.data
        K = a.shape[-1]
        pw = gemm.pack_weight(weight, cache=getattr(weight, "_sdf_cacheable", None))
        y, part = gemm.spike_gemm_fwd(a.view(-1, K), pw, None if bias is None else bias.detach(), want_stats, a_max=1)
        ctx.save_for_backward(weight)
        ctx.holder, ctx.has_bias, ctx.wt = holder, bias is not None, pw.wt
        y = y.view(*a.shape[:-1], weight.shape[0])
        if part is not None:
            ctx.mark_non_differentiable(part)
        return y, part

    @staticmethod
    def backward(ctx, gy, _gp):
        if gy is None:
            return (None,) * 5
        (weight,) = ctx.saved_tensors
        holder = ctx.holder
        a = holder.data
        K = a.shape[-1]
        g2 = gy.reshape(-1, gy.shape[-1])
        if g2.stride(1) != 1 or g2.stride(0) % 4 != 0 or g2.data_ptr() % 16 != 0:
            g2 = g2.contiguous()
        gtok = gw = gb = None
        if ctx.needs_input_grad[0]:
            wt = ctx.wt if ctx.wt is not None else weight.detach().t().contiguous()
            holder.add_grad(gemm.gemm_tf32(g2, wt).view(a.shape))
            gtok = _zero_token(gy.device)
        want_gb = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            gw = gemm.spike_wgrad(g2, a.view(-1, K), s_max=1, want_db=want_gb)     # bias gradient from the same pass over g
            if want_gb:
                gw, gb = gw
        elif want_gb:
            gb = g2.sum(0)
        return gtok, gw, gb, None, None


def spike_linear(s, weight, bias=None, exact_input=True, stats=None):
    """F.linear(s, weight, bias) for a spike (or small-integer) operand s with fp32-grade results on
    tensor cores; exact_input=False (real-valued s) keeps the plain fp32 forward GEMM.
    s: a Spikes (1-byte spikes) runs on the library's own tcgen05 GEMM; an fp32 tensor goes through the TF32 x 2 library
    path.  stats (bool, optional): when given the result is (y, partials) — the BN partial sums of y from the GEMM
    epilogue if stats is True and the tcgen05 path ran, else None."""
    if isinstance(s, Spikes):
        Cout = weight.shape[0]
        if Cout % 4:
            # the 2-channel flow heads (reference Spiking_modules.py:607-647): the engine's output rows are 16-byte pitched, so
            # the weight gets zero rows up to a multiple of 4 and the result is a view of the first Cout columns; autograd
            # pads / slices the (tiny) gradients accordingly
            padn = 4 - Cout % 4
            wp = torch.nn.functional.pad(weight, (0, 0, 0, padn))
            bp = None if bias is None else torch.nn.functional.pad(bias, (0, padn))
            y, part = _SpikeGemmFn.apply(s.token, wp, bp, s, bool(stats))
            y = y[..., :Cout]
            part = None if part is None else part[..., :Cout].contiguous()
        else:
            y, part = _SpikeGemmFn.apply(s.token, weight, bias, s, bool(stats))
        return y if stats is None else (y, part)
    if stats is not None:
        return spike_linear(s, weight, bias, exact_input), None
    if GEMM_MODE == "fp32":
        with _tf32(False):
            return torch.nn.functional.linear(s, weight, bias)
    if not exact_input:
        return _PlainLinearFn.apply(s, weight, bias)
    return _SpikeLinearFn.apply(s, weight, bias)


class _SpikeConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, transposed, output_padding):
        conv = torch.nn.functional.conv_transpose2d if transposed else torch.nn.functional.conv2d
        kw = dict(stride=stride, padding=padding)
        if transposed:
            kw["output_padding"] = output_padding
        hi, lo = split_tf32(weight.detach())
        with _tf32(True):
            y = conv(x, hi, None if bias is None else bias.detach(), **kw)
            y += conv(x, lo, None, **kw)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, padding, transposed, output_padding, bias is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        stride, padding, transposed, output_padding, has_bias = ctx.cfg
        two = lambda v: [v, v] if isinstance(v, int) else list(v)  # noqa: E731
        # keep the incoming gradient in the activation's memory format (NHWC): a plain .contiguous() here would be a
        # full NHWC->NCHW transpose of every conv gradient
        cl = x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        g = g.contiguous(memory_format=torch.channels_last) if cl else g.contiguous()
        with _tf32(True):
            gx, gw, gb = torch.ops.aten.convolution_backward(
                g, x, weight, [weight.shape[1] if transposed else weight.shape[0]] if has_bias else None,
                two(stride), two(padding), [1, 1], transposed, two(output_padding), 1,
                [ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]])
        return gx, gw, gb, None, None, None, None


class _PlainConvFn(torch.autograd.Function):
    """Real-valued operand: true fp32 forward, single-pass TF32 backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, transposed, output_padding):
        conv = torch.nn.functional.conv_transpose2d if transposed else torch.nn.functional.conv2d
        kw = dict(stride=stride, padding=padding)
        if transposed:
            kw["output_padding"] = output_padding
        with _tf32(False):
            y = conv(x, weight, bias, **kw)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, padding, transposed, output_padding, bias is not None)
        return y

    backward = _SpikeConvFn.backward


class _SmallCinConvFn(torch.autograd.Function):
    """3x3 / stride 1 / pad 1 conv with Cin <= 4 on a channels-last (N, H, W, Cin) tensor: direct fp32 kernel forward
    (patch-embed head, reference Spiking_modules.py:1737-1745) and direct fp32 weight / bias gradient kernel; the library
    (TF32) only when the input itself needs a gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _need_cuda(x, weight)
        x = x.contiguous()
        N, H, W, Cin = x.shape
        Cout = weight.shape[0]
        y = torch.empty((N, H, W, Cout), device=x.device, dtype=torch.float32)
        w = weight.detach().contiguous()
        capi.call("sdf_conv3x3_cl_fwd", capi.struct(
            "sdf_conv3x3_cl_args", x=_ptr(x), w=_ptr(w), bias=_ptr(None if bias is None else bias.detach().contiguous()),
            y=_ptr(y), N=N, H=H, W=W, Cin=Cin, Cout=Cout, stream=_stream()), algo_bytes=4 * (x.numel() + y.numel()))
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        if not ctx.needs_input_grad[0] and weight.shape[0] <= 128:
            # the model's case (the voxel input needs no gradient): own direct kernel for dW and db, true fp32
            g = g.contiguous()
            N, H, W, Cin = x.shape
            Cout = weight.shape[0]
            gw = torch.empty_like(weight, memory_format=torch.contiguous_format)
            gb = torch.empty(Cout, device=g.device, dtype=torch.float32) if ctx.has_bias else None
            nbytes = int(capi.lib().sdf_conv3x3_cl_wgrad_workspace_bytes(Cin, Cout))
            ws = torch.empty(nbytes // 4, device=g.device, dtype=torch.float32)
            capi.call("sdf_conv3x3_cl_wgrad", capi.struct(
                "sdf_conv3x3_cl_wgrad_args", x=_ptr(x), g=_ptr(g), dw=_ptr(gw), db=_ptr(gb), workspace=_ptr(ws),
                workspace_bytes=nbytes, N=N, H=H, W=W, Cin=Cin, Cout=Cout, stream=_stream()),
                algo_bytes=4 * (x.numel() + g.numel()))
            return None, (gw if ctx.needs_input_grad[1] else None), (gb if ctx.has_bias and ctx.needs_input_grad[2] else None)
        g4 = g.contiguous().permute(0, 3, 1, 2)           # logical NCHW, channels_last strides
        x4 = x.permute(0, 3, 1, 2)
        with _tf32(True):
            gx, gw, gb = torch.ops.aten.convolution_backward(
                g4, x4, weight, [weight.shape[0]] if ctx.has_bias else None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                [ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]])
        if gx is not None:
            gx = gx.permute(0, 2, 3, 1)
        return gx, gw, gb


def conv3x3_small_cin(x, weight, bias=None):
    return _SmallCinConvFn.apply(x, weight, bias)


class _SpikeConvGemmFn(torch.autograd.Function):
    """NHWC convolution of 1-byte spikes as an implicit GEMM on the tcgen05 engine (gemm.spike_conv_fwd: per-tap TMA boxes,
    zero padding = TMA out-of-bounds fill), weight gradient (gemm.spike_conv_wgrad) and stride-1 data gradient
    (gemm.conv_dgrad_tf32) and 3x3 / stride-2 data gradient (gemm.conv_dgrad_s2_tf32) on the same engine.  Reference: sj_layer.Conv2d on spike tensors, Spiking_modules.py:268,318,803,845-846."""

    @staticmethod
    def forward(ctx, token, weight, bias, holder, stride, padding, want_stats):
        ctx.set_materialize_grads(False)      # the BN partial sums get no gradient: do not zero-fill one per call
        x = holder.data                                   # (..., H, W, Cin) u8, leading dims = images
        H, W, Cin = x.shape[-3:]
        kh, kw = weight.shape[2], weight.shape[3]
        pw = gemm.pack_weight(weight, "conv")
        y, part = gemm.spike_conv_fwd(x.view(-1, H, W, Cin), pw, None if bias is None else bias.detach(), kh, kw, stride,
                                      padding, want_stats, a_max=1)
        ctx.save_for_backward(weight)
        ctx.holder, ctx.cfg, ctx.wt = holder, (stride, padding, bias is not None), pw.wt
        y = y.view(*x.shape[:-3], *y.shape[1:])
        if part is not None:
            ctx.mark_non_differentiable(part)
        return y, part

    @staticmethod
    def backward(ctx, gy, _gp):
        if gy is None:
            return (None,) * 7
        (weight,) = ctx.saved_tensors
        stride, padding, has_bias = ctx.cfg
        holder = ctx.holder
        x = holder.data
        H, W, Cin = x.shape[-3:]
        kh, kw = weight.shape[2], weight.shape[3]
        g4 = gy.contiguous().view(-1, *gy.shape[-3:])      # (Nimg, Ho, Wo, Cout) NHWC
        gtok = gw = gb = None
        if ctx.needs_input_grad[0]:
            if stride == 1 and weight.shape[0] % 32 == 0:
                gx = gemm.conv_dgrad_tf32(g4, weight, H, W, padding, ctx.wt)      # own implicit GEMM (TF32)
            elif stride == 2 and (kh, kw, padding) == (3, 3, 1) and weight.shape[0] % 32 == 0 and Cin % 4 == 0:
                gx = gemm.conv_dgrad_s2_tf32(g4, weight, H, W, ctx.wt)            # four parity-class launches of the same GEMM
            else:
                # other geometries: cuDNN; only the input SIZE and layout matter for a data gradient (channels-last,
                # so that cuDNN neither transposes g nor returns an NCHW result)
                fake = g4.new_empty((g4.shape[0], H, W, Cin)).permute(0, 3, 1, 2)
                with _tf32(True):
                    gx = torch.ops.aten.convolution_backward(
                        g4.permute(0, 3, 1, 2), fake, weight, None, [stride, stride], [padding, padding], [1, 1], False, [0, 0],
                        1, [True, False, False])[0]
                gx = gx.permute(0, 2, 3, 1).contiguous()      # no-op when cuDNN returned NHWC
            holder.add_grad(gx.view(x.shape))
            gtok = _zero_token(gy.device)
        want_gb = has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            gw = gemm.spike_conv_wgrad(g4, x.view(-1, H, W, Cin), kh, kw, stride, padding, s_max=1, want_db=want_gb)
            if want_gb:
                gw, gb = gw
        elif want_gb:
            gb = g4.sum((0, 1, 2))
        return gtok, gw, gb, None, None, None, None


class _SpikeDeconvFn(torch.autograd.Function):
    """ConvTranspose2d(3, stride 2, padding 1, output_padding 1) of 1-byte spikes on the tcgen05 engine (gemm.spike_deconv_fwd:
    four parity-class implicit GEMMs, exact integer contraction, BN sums from the epilogue).  Backward on the same engine:
    data gradient = a stride-2 TF32 convolution of g (gemm.deconv_dgrad_tf32), weight + bias gradient = four parity-class
    launches of G3 over strided views of g (gemm.spike_deconv_wgrad).  Reference: SpikingTransposeDecoderLayer.deconv,
    Spiking_modules.py:398-474."""

    @staticmethod
    def forward(ctx, token, weight, bias, holder, want_stats):
        ctx.set_materialize_grads(False)      # the BN partial sums get no gradient: do not zero-fill one per call
        x = holder.data                                   # (..., H, W, Cin_padded) u8
        H, W, Cin = x.shape[-3:]
        packs = gemm.pack_deconv_weight(weight, cin=Cin)
        y, part = gemm.spike_deconv_fwd(x.view(-1, H, W, Cin), packs, None if bias is None else bias.detach(), want_stats, a_max=1)
        ctx.save_for_backward(weight)
        ctx.holder, ctx.has_bias = holder, bias is not None
        y = y.view(*x.shape[:-3], 2 * H, 2 * W, weight.shape[1])
        if part is not None:
            ctx.mark_non_differentiable(part)
        return y, part

    @staticmethod
    def backward(ctx, gy, _gp):
        if gy is None:
            return (None,) * 5
        (weight,) = ctx.saved_tensors
        holder = ctx.holder
        x = holder.data
        H, W, Cin = x.shape[-3:]
        if DECONV_BWD == "lib":
            return _SpikeDeconvFn._backward_lib(ctx, gy)
        g4 = gy.contiguous().view(-1, *gy.shape[-3:])      # (Nimg, 2H, 2W, Cout) NHWC
        gtok = gw = gb = None
        if ctx.needs_input_grad[0]:
            # a stride-2 convolution of g: one TF32 implicit GEMM reading g through an element-stride-2 TMA map; the
            # gradient of the zero channels appended to the operand comes out as zeros
            holder.add_grad(gemm.deconv_dgrad_tf32(g4, weight, Cin=Cin).view(x.shape))
            gtok = _zero_token(gy.device)
        want_gb = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            gw = gemm.spike_deconv_wgrad(g4, x.view(-1, H, W, Cin), Cin_w=weight.shape[0], s_max=1, want_db=want_gb)
            if want_gb:
                gw, gb = gw
        elif want_gb:
            gb = g4.sum((0, 1, 2))
        return gtok, gw, gb, None, None


    @staticmethod
    def _backward_lib(ctx, gy):
        """Comparison path (SDF_DECONV_BWD=lib): cuDNN (TF32) on the spikes expanded to fp32 — what tools/bench_conv_bwd.py
        times the engine's kernels against."""
        (weight,) = ctx.saved_tensors
        holder = ctx.holder
        x = holder.data
        H, W, Cin = x.shape[-3:]
        g4 = gy.contiguous().view(-1, *gy.shape[-3:]).permute(0, 3, 1, 2)      # logical NCHW, channels-last strides
        x4 = x.view(-1, H, W, Cin).float().permute(0, 3, 1, 2)
        w = weight.detach()
        if Cin != w.shape[0]:
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, Cin - w.shape[0]))
        with _tf32(True):
            gx, gw, gb = torch.ops.aten.convolution_backward(
                g4, x4, w, [w.shape[1]] if ctx.has_bias else None, [2, 2], [1, 1], [1, 1], True, [1, 1], 1,
                [ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]])
        gtok = None
        if gx is not None:
            holder.add_grad(gx.permute(0, 2, 3, 1).contiguous().view(x.shape))
            gtok = _zero_token(gy.device)
        if gw is not None and Cin != weight.shape[0]:
            gw = gw[:weight.shape[0]]
        return gtok, gw, gb, None, None


USE_SPIKE_DECONV = True      # False: transposed convolutions stay on the library path (debug switch)
DECONV_BWD = os.environ.get("SDF_DECONV_BWD", "own")     # "lib": backward of the transposed convolutions through cuDNN (comparison)


def spike_deconv_supported(conv, cin):
    """Geometries gemm.spike_deconv_fwd takes: the x2 decoder up-sampling of every shipped model."""
    return (USE_SPIKE_DECONV and isinstance(conv, torch.nn.ConvTranspose2d) and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (2, 2)
            and tuple(conv.padding) == (1, 1) and tuple(conv.output_padding) == (1, 1) and conv.groups == 1
            and tuple(conv.dilation) == (1, 1) and cin % 16 == 0 and cin >= conv.in_channels and conv.out_channels % 4 == 0)


def spike_deconv(s: "Spikes", weight, bias=None, stats=None):
    """conv_transpose2d (3, stride 2, padding 1, output_padding 1) on a Spikes tensor (..., H, W, Cin) -> fp32 (..., 2H, 2W, Cout);
    `stats` as in spike_linear.  s may carry more channels than the weight (zero channels appended by the caller)."""
    y, part = _SpikeDeconvFn.apply(s.token, weight, bias, s, bool(stats))
    return y if stats is None else (y, part)


def spike_conv_supported(Cin, Cout, kernel_size, stride, padding):
    """Geometries the tcgen05 implicit-GEMM convolution takes (everything the SNN conv stack uses except transposed
    convolutions and the 2-channel heads)."""
    kh, kw = kernel_size
    return (Cin % 16 == 0 and Cout % 4 == 0 and kh * kw <= 9 and tuple(stride) in ((1, 1), (2, 2))
            and padding[0] == padding[1] and (Cin <= 256 or Cin % 256 == 0))


def spike_conv_gemm(s: "Spikes", weight, bias=None, stride=1, padding=0, stats=None):
    """conv2d on a Spikes tensor (..., H, W, Cin) -> fp32 (..., Ho, Wo, Cout) channels-last; `stats` as in spike_linear."""
    y, part = _SpikeConvGemmFn.apply(s.token, weight, bias, s, stride, padding, bool(stats))
    return y if stats is None else (y, part)


def spike_conv2d(x, weight, bias=None, stride=1, padding=0, transposed=False, output_padding=0, exact_input=True):
    """conv2d / conv_transpose2d on a spike operand, fp32-grade on tensor cores (see split_tf32)."""
    if GEMM_MODE != "fp32" and not exact_input:
        return _PlainConvFn.apply(x, weight, bias, stride, padding, transposed, output_padding)
    if GEMM_MODE == "fp32" or not exact_input:
        with _tf32(False):
            if transposed:
                return torch.nn.functional.conv_transpose2d(x, weight, bias, stride=stride, padding=padding,
                                                            output_padding=output_padding)
            return torch.nn.functional.conv2d(x, weight, bias, stride=stride, padding=padding)
    return _SpikeConvFn.apply(x, weight, bias, stride, padding, transposed, output_padding)


# ---------------------------------------------------------------------------------------------
# input pipeline
# ---------------------------------------------------------------------------------------------
def prepare_voxels(x, num_steps, normalize=True):
    """Signed voxel grid (B, bins, H, W) -> channels-last network input (B, steps, H, W, 2*bins/steps): polarity split,
    min-max normalisation over the non-zero entries and the bins -> (steps, channels) regroup of the patch embedding in one
    kernel pair (reference train_flow_parallel_supervised_SNN.py:261-265,278-284 + Spiking_modules.py:1772-1786).
    No autograd (it is the data side of the model)."""
    _need_cuda(x)
    x = x.detach().contiguous()
    B, bins, Hh, Ww = x.shape
    out = torch.empty((B, num_steps, Hh, Ww, 2 * bins // num_steps), device=x.device, dtype=torch.float32)
    ws = torch.empty(888, device=x.device, dtype=torch.float32) if normalize else None
    capi.call("sdf_voxel_prepare", capi.struct(
        "sdf_voxel_prepare_args", x=_ptr(x), out=_ptr(out), workspace=_ptr(ws), B=B, bins=bins, H=Hh, W=Ww, steps=num_steps,
        split=1, normalize=1 if normalize else 0, stream=_stream()), algo_bytes=(8 if normalize else 4) * x.numel() + 4 * out.numel())
    return out


def regroup_voxels(x, num_steps):
    """(B, bins, 2, H, W) — what the reference scripts hand to the model — -> (B, steps, H, W, 2*bins/steps) channels-last
    (the regroup of MS_PED_Spiking_PatchEmbed_Conv_sfn.forward, Spiking_modules.py:1772-1786, without the permute copy)."""
    _need_cuda(x)
    x = x.detach().contiguous()
    B, bins, two, Hh, Ww = x.shape
    assert two == 2
    out = torch.empty((B, num_steps, Hh, Ww, 2 * bins // num_steps), device=x.device, dtype=torch.float32)
    capi.call("sdf_voxel_prepare", capi.struct(
        "sdf_voxel_prepare_args", x=_ptr(x), out=_ptr(out), B=B, bins=bins, H=Hh, W=Ww, steps=num_steps, split=0, normalize=0,
        stream=_stream()), algo_bytes=4 * (x.numel() + out.numel()))
    return out


# ---------------------------------------------------------------------------------------------
# window index algebra
# ---------------------------------------------------------------------------------------------
class WindowGeom:
    """Geometry of one (feature map, window, shift) combination and its cached index tables.
    get_window_size clamping (reference swin_transformer3D_v2.py:68-81) is applied here."""
    _cache = {}

    def __init__(self, B, D, H, W, window, shift):
        ws, ss = list(window), list(shift)
        for i, x in enumerate((D, H, W)):
            if x <= ws[i]:
                ws[i] = x
                ss[i] = 0
        self.B, self.D, self.H, self.W = B, D, H, W
        self.window, self.shift = tuple(ws), tuple(ss)
        wd, wh, ww = self.window
        self.nD, self.nHw, self.nWw = -(-D // wd), -(-H // wh), -(-W // ww)
        self.Dp, self.Hp, self.Wp = self.nD * wd, self.nHw * wh, self.nWw * ww
        self.P = wh * ww
        self.N = wd * self.P
        self.nW = self.nD * self.nHw * self.nWw
        self.M = B * self.nW
        self.rows = self.M * self.N
        self.shifted = any(s > 0 for s in self.shift)
        self.win2x = None
        self.region = None

    def c(self):
        return dict(B=self.B, D=self.D, H=self.H, W=self.W, wd=self.window[0], wh=self.window[1], ww=self.window[2],
                    sd=self.shift[0], sh=self.shift[1], sw=self.shift[2])

    @classmethod
    def get(cls, B, D, H, W, window, shift, device):
        key = (B, D, H, W, tuple(window), tuple(shift), str(device))
        g = cls._cache.get(key)
        if g is None:
            g = cls(B, D, H, W, window, shift)
            g.build(device)
            cls._cache[key] = g
        return g

    def build(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("sdformerflow_b200: window tables are built on the GPU (no CPU fallback)")
        self.win2x = torch.empty(self.rows, device=device, dtype=torch.int32)
        self.region = torch.empty(self.nW * self.N, device=device, dtype=torch.uint8)
        capi.call("sdf_window_index", capi.struct("sdf_window_index_args", g=self.c(), win2x=_ptr(self.win2x),
                                                  region=_ptr(self.region), stream=_stream()))
        return self


class _WindowGatherFn(torch.autograd.Function):
    """x_windows = window_partition_v2(roll(pad(x))) (reference Spiking_swin_transformer3D.py:793-804)."""

    @staticmethod
    def forward(ctx, x, geom):
        _need_cuda(x)
        x = x.contiguous()
        C = x.shape[-1]
        wd, wh, ww = geom.window
        xw = torch.empty((wd, geom.M, wh, ww, C), device=x.device, dtype=torch.float32)
        capi.call("sdf_window_gather", capi.struct("sdf_window_gather_args", x=_ptr(x), xw=_ptr(xw),
                                                   win2x=_ptr(geom.win2x), rows=geom.rows, C=C, stream=_stream()))
        ctx.geom, ctx.xshape = geom, x.shape
        return xw

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        geom = ctx.geom
        dx = torch.empty(ctx.xshape, device=g.device, dtype=torch.float32)
        capi.call("sdf_window_gather_bwd", capi.struct("sdf_window_gather_bwd_args", dxw=_ptr(g), dx=_ptr(dx),
                                                       win2x=_ptr(geom.win2x), rows=geom.rows, C=ctx.xshape[-1],
                                                       stream=_stream()))
        return dx, None


def window_gather(x, geom):
    return _WindowGatherFn.apply(x, geom)


class _LifWindowFn(torch.autograd.Function):
    """proj_sn(x_windows) with the pad/roll/partition gather folded in (reference :670 / :425)."""

    @staticmethod
    def forward(ctx, x, geom, cfg, holder=None):
        _need_cuda(x)
        x = x.contiguous()
        C = x.shape[-1]
        wd, wh, ww = geom.window
        dt = capi.SDF_SPIKE_F32 if holder is None else capi.SDF_SPIKE_U8
        spike = torch.empty((wd, geom.M, wh, ww, C), device=x.device, dtype=_SPIKE_TORCH[dt])
        cc = cfg.c()
        h_tap = torch.empty(spike.shape, device=x.device, dtype=torch.float32) if _tapping() else None
        capi.call("sdf_lif_window_fwd", capi.struct(
            "sdf_lif_window_fwd_args", x=_ptr(x), spike=_ptr(spike), h_seq=_ptr(h_tap), win2x=_ptr(geom.win2x), wd=wd,
            MP=geom.M * geom.P, C=C, neuron=cc, spike_dtype=dt, stream=_stream()),
            algo_bytes=4 * x.numel() + spike.element_size() * spike.numel())
        _tap_take(1, h_tap)
        ctx.save_for_backward(x)
        ctx.geom, ctx.cc, ctx.holder = geom, cc, holder
        if holder is not None:
            holder.data = spike
            return x.new_empty(())
        return spike

    @staticmethod
    def backward(ctx, gs):
        (x,) = ctx.saved_tensors
        gs = gs.contiguous() if ctx.holder is None else ctx.holder.take_grad()
        geom = ctx.geom
        gx = torch.empty_like(x)
        capi.call("sdf_lif_window_bwd", capi.struct(
            "sdf_lif_window_bwd_args", x=_ptr(x), grad_spike=_ptr(gs), grad_x=_ptr(gx), win2x=_ptr(geom.win2x),
            wd=geom.window[0], MP=geom.M * geom.P, C=x.shape[-1], neuron=ctx.cc, stream=_stream()))
        return gx, None, None, None


def lif_window(x, geom, cfg: NeuronCfg, u8=False):
    if cfg.kind == capi.SDF_NEURON_PLIF:
        raise NotImplementedError("lif_window: ParametricLIFNode is not fused here (use window_gather + neuron)")
    if not u8:
        return _LifWindowFn.apply(x, geom, cfg, None)
    holder = Spikes()
    holder.token = _LifWindowFn.apply(x, geom, cfg, holder)
    return holder


def lif_window_debug(x, geom, cfg: NeuronCfg):
    x = x.contiguous()
    C = x.shape[-1]
    wd, wh, ww = geom.window
    spike = torch.empty((wd, geom.M, wh, ww, C), device=x.device, dtype=torch.float32)
    h = torch.empty_like(spike)
    capi.call("sdf_lif_window_fwd", capi.struct(
        "sdf_lif_window_fwd_args", x=_ptr(x), spike=_ptr(spike), h_seq=_ptr(h), win2x=_ptr(geom.win2x), wd=wd,
        MP=geom.M * geom.P, C=C, neuron=cfg.c(), spike_dtype=capi.SDF_SPIKE_F32, stream=_stream()))
    return spike, h


class _WindowScatterFn(torch.autograd.Function):
    """out = shortcut + alpha_b * BN(y)[window -> token]: proj_bn + window_reverse + roll back + crop +
    DropPath + residual (reference Spiking_swin_transformer3D.py:713-715, :810-820, :840)."""

    @staticmethod
    def forward(ctx, y, res, weight, bias, bn, geom, alpha, partials=None):
        _need_cuda(y, res)
        y = y.contiguous()
        C = y.shape[-1]
        rows = geom.rows
        assert y.numel() == rows * C
        scale = shift = mean = rstd = None
        if bn is not None:
            scale, shift, mean, rstd = _bn_forward_affine(y, rows, C, C, bn, partials)
        out = torch.empty((geom.B, geom.D, geom.H, geom.W, C), device=y.device, dtype=torch.float32)
        if res is not None:
            res = res.contiguous()
        capi.call("sdf_window_scatter", capi.struct(
            "sdf_window_scatter_args", y=_ptr(y), res=_ptr(res), out=_ptr(out), win2x=_ptr(geom.win2x),
            scale=_ptr(scale), shift=_ptr(shift), alpha=_ptr(alpha), rows=rows, rows_per_sample=geom.nW * geom.N, C=C,
            stream=_stream()), algo_bytes=4 * rows * C + (4 if res is None else 8) * out.numel())
        ctx.save_for_backward(y, weight, mean, rstd, alpha)
        ctx.geom, ctx.C, ctx.has_bn, ctx.training = geom, C, bn is not None, (bn.training if bn is not None else False)
        ctx.has_res = res is not None
        return out

    @staticmethod
    def backward(ctx, go):
        y, weight, mean, rstd, alpha = ctx.saved_tensors
        go = go.contiguous()
        geom, C = ctx.geom, ctx.C
        rows = geom.rows
        dy = torch.empty((rows, C), device=go.device, dtype=torch.float32)
        partials = torch.empty((N_PARTIAL, 2, C), device=go.device, dtype=torch.float32) if ctx.has_bn else None
        capi.call("sdf_window_scatter_bwd", capi.struct(
            "sdf_window_scatter_bwd_args", dout=_ptr(go), dy=_ptr(dy), u=_ptr(y) if ctx.has_bn else None,
            win2x=_ptr(geom.win2x), alpha=_ptr(alpha), bn_partials=_ptr(partials), n_partial_blocks=N_PARTIAL,
            rows=rows, rows_per_sample=geom.nW * geom.N, C=C, stream=_stream()),
            algo_bytes=4 * go.numel() + (8 if ctx.has_bn else 4) * rows * C)
        gw = gb = None
        if ctx.has_bn:
            dy, gw, gb = _bn_backward(partials, dy, y, C, rows, C, weight, mean, rstd, ctx.training)
        return dy.view(y.shape), (go if ctx.has_res else None), gw, gb, None, None, None, None


def window_scatter(y, geom, res=None, bn_module=None, alpha=None, partials=None):
    bn = BNParams(bn_module) if bn_module is not None else None
    w = bn.weight if bn is not None else None
    b = bn.bias if bn is not None else None
    return _WindowScatterFn.apply(y, res, w, b, bn, geom, alpha, partials)


# ---------------------------------------------------------------------------------------------
# patch merging gather (+ LIF over D)
# ---------------------------------------------------------------------------------------------
class _LifMergeFn(torch.autograd.Function):
    """MS_SpikingPatchMerging: sn(cat(x0..x3)) (reference :958-970); apply_neuron=False gives the
    plain 2x2 gather of SpikingPatchMerging (:919-930)."""

    @staticmethod
    def forward(ctx, x, cfg, apply_neuron, holder=None):
        _need_cuda(x)
        x = x.contiguous()
        B, D, H, W, C = x.shape
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        dt = capi.SDF_SPIKE_F32 if holder is None else capi.SDF_SPIKE_U8
        out = torch.empty((B, D, H2, W2, 4 * C), device=x.device, dtype=_SPIKE_TORCH[dt])
        cc = cfg.c() if apply_neuron else NeuronCfg().c()
        h_tap = torch.empty(out.shape, device=x.device, dtype=torch.float32) if (_tapping() and apply_neuron) else None
        capi.call("sdf_lif_merge_fwd", capi.struct(
            "sdf_lif_merge_fwd_args", x=_ptr(x), spike=_ptr(out), h_seq=_ptr(h_tap), B=B, D=D, H=H, W=W, C=C, neuron=cc,
            spike_dtype=dt, apply_neuron=1 if apply_neuron else 0, stream=_stream()))
        if apply_neuron:
            _tap_take(1, h_tap)
        ctx.save_for_backward(x)
        ctx.cc, ctx.apply_neuron, ctx.holder = cc, apply_neuron, holder
        if holder is not None:
            holder.data = out
            return x.new_empty(())
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.contiguous() if ctx.holder is None else ctx.holder.take_grad()
        B, D, H, W, C = x.shape
        gx = torch.empty_like(x)
        capi.call("sdf_lif_merge_bwd", capi.struct(
            "sdf_lif_merge_bwd_args", x=_ptr(x), grad_spike=_ptr(g), grad_x=_ptr(gx), B=B, D=D, H=H, W=W, C=C,
            neuron=ctx.cc, apply_neuron=1 if ctx.apply_neuron else 0, stream=_stream()))
        return gx, None, None, None


def lif_merge(x, cfg: NeuronCfg, apply_neuron=True, u8=False):
    if apply_neuron and cfg.kind == capi.SDF_NEURON_PLIF:
        raise NotImplementedError("lif_merge: ParametricLIFNode is not fused here (gather with apply_neuron=False + neuron)")
    if not u8:
        return _LifMergeFn.apply(x, cfg, apply_neuron, None)
    if not apply_neuron:
        raise ValueError("lif_merge(u8=True) needs apply_neuron=True (a plain gather of membranes is not a spike tensor)")
    holder = Spikes()
    holder.token = _LifMergeFn.apply(x, cfg, apply_neuron, holder)
    return holder


# ---------------------------------------------------------------------------------------------
# K5: QK token-gate attention core
# ---------------------------------------------------------------------------------------------
class _QKGateFn(torch.autograd.Function):
    """g = sn_k(bn_k(k_pre)+pos) * sn2_q(sum_d sn_q(bn_q(q_pre))), permuted to proj-input order
    (reference Spiking_swin_transformer3D.py:671-710).  qk_pre: [wd*M*P, 2C] = [q_pre | k_pre]."""

    @staticmethod
    def forward(ctx, qk_pre, wq, bq, wk, bk, pos, bn_q, bn_k, cfg, wd, M, P, nH, part_q=None, part_k=None, holder=None):
        _need_cuda(qk_pre, pos)
        qk_pre = qk_pre.contiguous()
        rows = wd * M * P
        C = nH * 32
        assert qk_pre.shape == (rows, 2 * C)
        q_pre, k_pre = qk_pre[:, :C], qk_pre[:, C:]
        qs, qh, qm, qr = _bn_forward_affine(q_pre, rows, C, 2 * C, bn_q, part_q)
        ks, kh, km, kr = _bn_forward_affine(k_pre, rows, C, 2 * C, bn_k, part_k)
        posc = pos.detach().contiguous()
        dt = capi.SDF_SPIKE_F32 if holder is None else capi.SDF_SPIKE_U8
        gate = torch.empty((rows, C), device=qk_pre.device, dtype=_SPIKE_TORCH[dt])
        cc = cfg.c()
        tq = tk = ta = None
        if _tapping():          # membranes of sn_q, sn_k, sn2_q (the order the module announces them in)
            tq = torch.empty((rows, C), device=qk_pre.device, dtype=torch.float32)
            tk = torch.empty_like(tq)
            ta = torch.empty((rows, nH), device=qk_pre.device, dtype=torch.float32)
        capi.call("sdf_attn_qkgate_fwd", capi.struct(
            "sdf_attn_qkgate_fwd_args", q_pre=_ptr(q_pre), k_pre=_ptr(k_pre), ld=2 * C, q_scale=_ptr(qs), q_shift=_ptr(qh),
            k_scale=_ptr(ks), k_shift=_ptr(kh), pos=_ptr(posc), gate=_ptr(gate), q_h=_ptr(tq), k_h=_ptr(tk), a_h=_ptr(ta),
            wd=wd, M=M, P=P, C=C, nH=nH, neuron=cc, spike_dtype=dt, stream=_stream()),
            algo_bytes=(8 + gate.element_size()) * rows * C)
        _tap_take(3, tq, tk, ta)
        ctx.save_for_backward(qk_pre, wq, wk, pos, qs, qh, qm, qr, ks, kh, km, kr)
        ctx.dims, ctx.cc, ctx.tq, ctx.tk, ctx.holder = (wd, M, P, C, nH), cc, bn_q.training, bn_k.training, holder
        if holder is not None:
            holder.data = gate
            return qk_pre.new_empty(())
        return gate

    @staticmethod
    def backward(ctx, gg):
        qk_pre, wq, wk, pos, qs, qh, qm, qr, ks, kh, km, kr = ctx.saved_tensors
        wd, M, P, C, nH = ctx.dims
        rows = wd * M * P
        gg = gg.contiguous() if ctx.holder is None else ctx.holder.take_grad()
        dev = gg.device
        q_pre, k_pre = qk_pre[:, :C], qk_pre[:, C:]
        gq = torch.empty((rows, C), device=dev, dtype=torch.float32)
        gk = torch.empty_like(gq)
        pq = torch.empty((N_PARTIAL, 2, C), device=dev, dtype=torch.float32)
        pk = torch.empty_like(pq)
        capi.call("sdf_attn_qkgate_bwd", capi.struct(
            "sdf_attn_qkgate_bwd_args", q_pre=_ptr(q_pre), k_pre=_ptr(k_pre), ld=2 * C, q_scale=_ptr(qs), q_shift=_ptr(qh),
            k_scale=_ptr(ks), k_shift=_ptr(kh), pos=_ptr(pos.detach().contiguous()), grad_gate=_ptr(gg), grad_q=_ptr(gq),
            grad_k=_ptr(gk), bn_partials_q=_ptr(pq), bn_partials_k=_ptr(pk), n_partial_blocks=N_PARTIAL, wd=wd, M=M,
            P=P, C=C, nH=nH, neuron=ctx.cc, stream=_stream()), algo_bytes=20 * rows * C)
        gpos = torch.empty(pos.numel(), device=dev, dtype=torch.float32)
        capi.call("sdf_pos_grad", capi.struct("sdf_pos_grad_args", grad_k=_ptr(gk), grad_pos=_ptr(gpos), wd=wd, M=M,
                                              PC=P * C, stream=_stream()))
        # BN backward of both halves, written into one [rows, 2C] buffer for the shared GEMM backward
        dqk = torch.empty((rows, 2 * C), device=dev, dtype=torch.float32)
        outs = []
        for half, (g, u, w, m, r, tr) in enumerate(((gq, q_pre, wq, qm, qr, ctx.tq), (gk, k_pre, wk, km, kr, ctx.tk))):
            gw = torch.empty(C, device=dev, dtype=torch.float32)
            gb = torch.empty_like(gw)
            coef = torch.empty((3, C), device=dev, dtype=torch.float32)
            capi.call("sdf_bn_bwd_finalize", capi.struct(
                "sdf_bn_bwd_finalize_args", partials=_ptr(pq if half == 0 else pk), n_partial_blocks=N_PARTIAL,
                count=rows, C=C, weight=_ptr(w), mean=_ptr(m), rstd=_ptr(r), grad_weight=_ptr(gw), grad_bias=_ptr(gb),
                coef=_ptr(coef), training=1 if tr else 0, stream=_stream()))
            capi.call("sdf_bn_bwd_apply", capi.struct(
                "sdf_bn_bwd_apply_args", dy=_ptr(g), u=_ptr(u), ld_u=2 * C, du=_ptr(dqk[:, half * C:]), ld_du=2 * C,
                coef=_ptr(coef), rows=rows, C=C, stream=_stream()), algo_bytes=12 * rows * C)
            outs += [gw, gb]
        return (dqk, outs[0], outs[1], outs[2], outs[3], gpos.view(pos.shape), None, None, None, None, None, None, None,
                None, None, None)


def qkgate(qk_pre, bn_q_module, bn_k_module, pos, cfg: NeuronCfg, wd, M, P, nH, partials=None, u8=False):
    """partials: BN partial sums [N_PARTIAL, 2, 2C] of qk_pre = [q_pre | k_pre] from the GEMM epilogue; u8: gate spikes as
    a Spikes for the proj GEMM."""
    if cfg.kind == capi.SDF_NEURON_PLIF:
        raise NotImplementedError("qkgate: ParametricLIFNode is not fused here (the module falls back to the generic path)")
    bq, bk = BNParams(bn_q_module), BNParams(bn_k_module)
    C = nH * 32
    pq = pk = None
    if partials is not None:
        pq, pk = partials[:, :, :C].contiguous(), partials[:, :, C:].contiguous()
    holder = Spikes() if u8 else None
    out = _QKGateFn.apply(qk_pre, bq.weight, bq.bias, bk.weight, bk.bias, pos, bq, bk, cfg, wd, M, P, nH, pq, pk, holder)
    if u8:
        holder.token = out
        return holder
    return out


def qkgate_debug(q_pre, k_pre, q_scale, q_shift, k_scale, k_shift, pos, cfg, wd, M, P, nH):
    """forward only, returning (gate, q membrane, k membrane, sn2_q membrane) for parity tests."""
    rows, C = wd * M * P, nH * 32
    dev = q_pre.device
    gate = torch.empty((rows, C), device=dev, dtype=torch.float32)
    qh = torch.empty_like(gate)
    kh = torch.empty_like(gate)
    ah = torch.empty((rows, nH), device=dev, dtype=torch.float32)
    capi.call("sdf_attn_qkgate_fwd", capi.struct(
        "sdf_attn_qkgate_fwd_args", q_pre=_ptr(q_pre), k_pre=_ptr(k_pre), ld=q_pre.stride(0), q_scale=_ptr(q_scale),
        q_shift=_ptr(q_shift), k_scale=_ptr(k_scale), k_shift=_ptr(k_shift), pos=_ptr(pos.contiguous()), gate=_ptr(gate),
        q_h=_ptr(qh), k_h=_ptr(kh), a_h=_ptr(ah), wd=wd, M=M, P=P, C=C, nH=nH, neuron=cfg.c(),
        spike_dtype=capi.SDF_SPIKE_F32, stream=_stream()))
    return gate, qh, kh, ah


# ---------------------------------------------------------------------------------------------
# K3/K4: Q K^T V attention on tcgen05 (with its three input neurons, so spikes stay 1 byte)
# ---------------------------------------------------------------------------------------------
def _neuron_u8(u, lay, cfg, scale, shift, C, psn):
    """spikes as uint8 {0,1} of neuron(BN(u)) — LIF family or PSN."""
    spike = torch.empty(u.shape, device=u.device, dtype=torch.uint8)
    if psn is None:
        capi.call("sdf_lif_fwd", capi.struct(
            "sdf_lif_fwd_args", u=_ptr(u), spike=_ptr(spike), scale=_ptr(scale), shift=_ptr(shift), C=C, hw=1, lay=lay,
            neuron=cfg.c(), spike_dtype=capi.SDF_SPIKE_U8, stream=_stream()), algo_bytes=5 * u.numel())
    else:
        capi.call("sdf_psn_fwd", capi.struct(
            "sdf_psn_fwd_args", u=_ptr(u), spike=_ptr(spike), weight=_ptr(psn[0]), bias=_ptr(psn[1]), scale=_ptr(scale),
            shift=_ptr(shift), C=C, hw=1, lay=lay, spike_dtype=capi.SDF_SPIKE_U8, stream=_stream()),
            algo_bytes=5 * u.numel())
    return spike


class _QKTVFn(torch.autograd.Function):
    """O = (scale * Q K^T + Bias + Mask) @ V with Q,K,V = sn(bn(x_pre)) (reference
    Spiking_swin_transformer3D.py:308-363); output rows in proj-input order."""

    @staticmethod
    def forward(ctx, q_pre, k_pre, v_pre, wq, bq, wk, bk, wv, bv, table, bns, cfg, region, dims, scale, want_attn,
                *psn_params):
        wd, M, P, nH, nW, ws = dims
        C, N = nH * 32, wd * P
        rows = wd * M * P
        _need_cuda(q_pre, k_pre, v_pre, table)
        pres = [t.contiguous() for t in (q_pre, k_pre, v_pre)]
        lay = seq_layout((wd, M * P * C), 0)
        spikes, saved = [], []
        for i, (u, bn) in enumerate(zip(pres, bns)):
            sc, sh, mean, rstd = _bn_forward_affine(u, rows, C, C, bn)
            psn = None if not psn_params else (psn_params[2 * i].detach().contiguous(), psn_params[2 * i + 1].detach().contiguous())
            spikes.append(_neuron_u8(u, lay, cfg, sc, sh, C, psn))
            saved += [sc, sh, mean, rstd]
        out = torch.empty((rows, C), device=q_pre.device, dtype=torch.float32)
        attn = torch.empty((M * nH, N, N), device=q_pre.device, dtype=torch.float32) if want_attn else None
        tab = table.detach().contiguous()
        capi.call("sdf_attn_qktv_fwd", capi.struct(
            "sdf_attn_qktv_fwd_args", q=_ptr(spikes[0]), k=_ptr(spikes[1]), v=_ptr(spikes[2]), bias_table=_ptr(tab),
            region=_ptr(region), out=_ptr(out), attn_dbg=_ptr(attn), M=M, nH=nH, nW=nW, wd=ws[0], wh=ws[1], ww=ws[2],
            scale=float(scale), stream=_stream()), algo_bytes=3 * rows * C + 4 * rows * C)
        ctx.save_for_backward(*pres, *spikes, wq, wk, wv, tab, *saved, *psn_params)
        ctx.meta = (dims, scale, cfg, region, lay, [b.training for b in bns], len(psn_params) > 0)
        if want_attn:
            ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, go, _ga):
        dims, scale, cfg, region, lay, trains, has_psn = ctx.meta
        wd, M, P, nH, nW, ws = dims
        C, N = nH * 32, wd * P
        rows = wd * M * P
        sv = ctx.saved_tensors
        pres, spikes, ws_bn, tab = sv[0:3], sv[3:6], sv[6:9], sv[9]
        stats = sv[10:22]
        psn_params = sv[22:] if has_psn else None
        dev = go.device
        go = go.contiguous()
        gq = torch.empty((rows, C), device=dev, dtype=torch.float32)
        gk, gv = torch.empty_like(gq), torch.empty_like(gq)
        gtab = torch.zeros_like(tab)
        capi.call("sdf_attn_qktv_bwd", capi.struct(
            "sdf_attn_qktv_bwd_args", q=_ptr(spikes[0]), k=_ptr(spikes[1]), v=_ptr(spikes[2]), bias_table=_ptr(tab),
            region=_ptr(region), grad_out=_ptr(go), grad_q=_ptr(gq), grad_k=_ptr(gk), grad_v=_ptr(gv),
            grad_bias_table=_ptr(gtab), M=M, nH=nH, nW=nW, wd=ws[0], wh=ws[1], ww=ws[2], scale=float(scale),
            stream=_stream()))
        outs, bn_grads, psn_grads = [], [], []
        for i, (u, g) in enumerate(zip(pres, (gq, gk, gv))):
            sc, sh, mean, rstd = stats[4 * i:4 * i + 4]
            partials = torch.empty((N_PARTIAL, 2, C), device=dev, dtype=torch.float32)
            dx = torch.empty_like(u)
            if not has_psn:
                capi.call("sdf_lif_bwd", capi.struct(
                    "sdf_lif_bwd_args", u=_ptr(u), grad_spike=_ptr(g), grad_x=_ptr(dx), scale=_ptr(sc), shift=_ptr(sh),
                    bn_partials=_ptr(partials), n_partial_blocks=N_PARTIAL, C=C, hw=1, lay=lay, neuron=cfg.c(),
                    stream=_stream()), algo_bytes=12 * u.numel())
            else:
                pw, pb = psn_params[2 * i].detach().contiguous(), psn_params[2 * i + 1].detach().contiguous()
                T, n = lay["T"], lay["n_neurons"]
                gh = torch.empty((T, n), device=dev, dtype=torch.float32)
                xo = torch.empty((T, n), device=dev, dtype=torch.float32)
                capi.call("sdf_psn_bwd", capi.struct(
                    "sdf_psn_bwd_args", u=_ptr(u), grad_spike=_ptr(g), grad_x=_ptr(dx), grad_h=_ptr(gh), x_out=_ptr(xo),
                    weight=_ptr(pw), bias=_ptr(pb), scale=_ptr(sc), shift=_ptr(sh), bn_partials=_ptr(partials),
                    n_partial_blocks=N_PARTIAL, C=C, hw=1, lay=lay, surrogate=cfg.surrogate, sg_alpha=float(cfg.sg_alpha),
                    stream=_stream()))
                psn_grads += list(_psn_param_grads(gh, xo))
            du, gw, gb = _bn_backward(partials, dx, u, C, rows, C, ws_bn[i], mean, rstd, trains[i])
            outs.append(du.view(u.shape))
            bn_grads += [gw, gb]
        return (outs[0], outs[1], outs[2], bn_grads[0], bn_grads[1], bn_grads[2], bn_grads[3], bn_grads[4], bn_grads[5],
                gtab, None, None, None, None, None, None, *psn_grads)


def qktv_attention(q_pre, k_pre, v_pre, bn_q, bn_k, bn_v, table, cfg, region, wd, M, P, nH, nW, window, scale,
                   psn_modules=None, want_attn=False):
    bns = [BNParams(b) for b in (bn_q, bn_k, bn_v)]
    dims = (wd, M, P, nH, nW, tuple(window))
    psn = [] if psn_modules is None else [t for m in psn_modules for t in (m.weight, m.bias)]
    return _QKTVFn.apply(q_pre, k_pre, v_pre, bns[0].weight, bns[0].bias, bns[1].weight, bns[1].bias, bns[2].weight,
                         bns[2].bias, table, bns, cfg, region, dims, scale, want_attn, *psn)


def qktv_debug(q, k, v, table, region, M, nH, nW, window, scale, debug=True):
    """Raw forward call on uint8 spikes [M*nH*N, 32]: returns (out [wd*M*P, C], S int32 [M*nH,N,N], attn fp32).
    debug=False leaves the two debug outputs out (returns None for them), which is what the model path does and
    what lets the library pick its fast kernel for small windows."""
    wd, wh, ww = window
    N, P = wd * wh * ww, wh * ww
    rows, C = wd * M * P, nH * 32
    dev = q.device
    out = torch.empty((rows, C), device=dev, dtype=torch.float32)
    s_dbg = torch.empty((M * nH, N, N), device=dev, dtype=torch.int32) if debug else None
    attn = torch.empty((M * nH, N, N), device=dev, dtype=torch.float32) if debug else None
    capi.call("sdf_attn_qktv_fwd", capi.struct(
        "sdf_attn_qktv_fwd_args", q=_ptr(q), k=_ptr(k), v=_ptr(v), bias_table=_ptr(table.contiguous()), region=_ptr(region),
        out=_ptr(out), s_dbg=_ptr(s_dbg), attn_dbg=_ptr(attn), M=M, nH=nH, nW=nW, wd=wd, wh=wh, ww=ww, scale=float(scale),
        stream=_stream()))
    return out, s_dbg, attn


def qktv_bwd_debug(q, k, v, table, region, grad_out, M, nH, nW, window, scale):
    wd, wh, ww = window
    N, P = wd * wh * ww, wh * ww
    rows, C = wd * M * P, nH * 32
    dev = q.device
    gq = torch.empty((rows, C), device=dev, dtype=torch.float32)
    gk, gv = torch.empty_like(gq), torch.empty_like(gq)
    gtab = torch.zeros_like(table)
    capi.call("sdf_attn_qktv_bwd", capi.struct(
        "sdf_attn_qktv_bwd_args", q=_ptr(q), k=_ptr(k), v=_ptr(v), bias_table=_ptr(table.contiguous()), region=_ptr(region),
        grad_out=_ptr(grad_out.contiguous()), grad_q=_ptr(gq), grad_k=_ptr(gk), grad_v=_ptr(gv), grad_bias_table=_ptr(gtab),
        M=M, nH=nH, nW=nW, wd=wd, wh=wh, ww=ww, scale=float(scale), stream=_stream()))
    return gq, gk, gv, gtab
