"""Raw host bindings of the tcgen05 + TMA GEMM engine (csrc/spike_gemm.cu, csrc/spike_wgrad.cu).

Forward: u8 spikes x 3 signed 8-bit weight digit planes (tcgen05.mma kind::i8, exact integer accumulate, one fp32 rounding).
Data gradient: fp32 x fp32 read as TF32.  Weight gradient: G^T S with a split over the row (pixel) axis.
Autograd wrappers live in ops.py; this module only allocates outputs and calls the C-ABI.
"""
import weakref

import torch

from . import capi

N_PARTIAL = 444  # capacity of every BN partial-sum workspace ([N_PARTIAL, 2, C]); == ops.N_PARTIAL


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class PackedWeight:
    """Digit planes + scales of one Linear / Conv weight, valid for one (data_ptr, _version) of the fp32 parameter."""
    __slots__ = ("wq", "wscale", "wt", "Cout", "Cin", "taps", "key", "owner")

    def __init__(self, wq, wscale, wt, Cout, Cin, taps, key, owner=None):
        self.wq, self.wscale, self.wt, self.Cout, self.Cin, self.taps, self.key = wq, wscale, wt, Cout, Cin, taps, key
        self.owner = owner          # weakref to the parameter this pack belongs to (None: not cached)


_pack_cache = {}

# Every cache of a weight-derived quantity (digit planes, concatenated q|k weight, PLIF 1/tau) is keyed by the parameter's
# (data_ptr, _version) AND by this epoch.  ``_version`` alone is not enough: fused / foreach optimizers update parameters
# without bumping it (torch 2.11 _fused_adamw_), so a global optimizer-step hook bumps the epoch after ANY optimizer.step(),
# and train.GraphedStep bumps it after every replay (an optimizer captured in a CUDA graph runs no Python hook).  Weights
# edited through .data / raw pointers, or by a user-made graph that contains the optimizer: call invalidate_pack_cache().
_epoch = [0]
_CACHE_OK = True


def bump_weights_epoch():
    _epoch[0] += 1


def weights_epoch():
    return _epoch[0]


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook
    _reg_hook(lambda _opt, _args, _kwargs: bump_weights_epoch())
except ImportError:          # no global hook on this torch: never cache
    _CACHE_OK = False


def _layout_dims(wd, layout):
    if layout == "linear":
        Cout, Cin = wd.shape
        return Cout, Cin, 1, Cin, 1, 0
    if layout == "conv":
        Cout, Cin, kh, kw = wd.shape
        taps = kh * kw
        return Cout, Cin, taps, Cin * taps, taps, 1
    raise ValueError(layout)


def _pack_launch(wd, layout, wq, wscale, wt):
    Cout, Cin, taps, s_co, s_ci, s_tap = _layout_dims(wd, layout)
    capi.call("sdf_spike_gemm_pack", capi.struct(
        "sdf_spike_gemm_pack_args", w=_ptr(wd), wq=_ptr(wq), wscale=_ptr(wscale), wt=_ptr(wt), wq_bytes=wq.numel(), Cout=Cout, Cin=Cin,
        taps=taps, s_co=s_co, s_ci=s_ci, s_tap=s_tap, tap_map=list(range(9)), stream=_stream()))


# ---- pre-packing plan (train.GraphedStep) ------------------------------------------------------------------------------
# A captured training step re-packs every weight on every replay (the optimizer has just changed them).  Issued lazily, layer
# by layer, those ~80 small kernels sit on the step's critical path one after the other; a PackPlan issues them all at the
# START of the step on side streams (forked from the capturing stream, so they become parallel branches of the graph) into
# persistent buffers, and the layers' pack_weight / pack_deconv_weight calls pick those buffers up (joining the branches at
# the first hit) — the packs overlap with each other and with the head of the forward pass, which does not need them.
_record = None      # list of jobs while start_recording() is active
_pinned = {}        # id(parameter) -> (job key, result, weakref, plan): results an active PackPlan owns


def start_recording():
    global _record
    _record = []


def stop_recording():
    global _record
    jobs, _record = _record or [], None
    seen, out = set(), []
    for job in jobs:
        k = (id(job[1]),) + tuple(job[2:])
        if k not in seen:
            seen.add(k)
            out.append(job)
    return out


def _pinned_hit(w, jobkey):
    hit = _pinned.get(id(w))
    if hit is None or hit[0] != jobkey or hit[2]() is not w:
        return None
    hit[3].join()
    return hit[1]


class PackPlan:
    """jobs: what stop_recording() returned for one forward pass of the model (parameters only)."""

    def __init__(self, jobs, n_streams=4):
        self.jobs = []
        for kind, w, *rest in jobs:
            if kind == "w":
                layout, need_wt = rest
                res = pack_weight(w, layout, cache=False, need_wt=need_wt)             # allocates the persistent buffers
            else:
                (cin,) = rest
                res = pack_deconv_weight(w, cin=cin, cache=False)
            self.jobs.append((kind, w, tuple(rest), res))
        self.streams = [torch.cuda.Stream() for _ in range(max(1, n_streams))]
        self.joined = True
        self._main = None

    def run(self):
        """Re-pack every weight into the plan's buffers on the side streams and pin the results for the layers."""
        self._main = torch.cuda.current_stream()
        for st in self.streams:
            st.wait_stream(self._main)
        for j, (kind, w, rest, res) in enumerate(self.jobs):
            with torch.cuda.stream(self.streams[j % len(self.streams)]):
                if kind == "w":
                    wd = w.detach()
                    _pack_launch(wd if wd.is_contiguous() else wd.contiguous(), rest[0], res.wq, res.wscale, res.wt)
                else:
                    _deconv_pack_launch(w, rest[0], res)
            _pinned[id(w)] = ((kind,) + rest, res, weakref.ref(w), self)
        self.joined = False

    def join(self):
        if not self.joined:
            cur = torch.cuda.current_stream()
            for st in self.streams:
                cur.wait_stream(st)
            self.joined = True

    def release(self):
        """Un-pin (the buffers stay allocated for the next run())."""
        self.join()
        for kind, w, rest, res in self.jobs:
            hit = _pinned.get(id(w))
            if hit is not None and hit[3] is self:
                del _pinned[id(w)]


def pack_weight(w, layout="linear", cache=None, need_wt=False):
    """fp32 weight -> PackedWeight.  layout: 'linear' (Cout, K), 'conv' (Cout, Cin, kh, kw) OIHW.
    Cached per parameter version (optimizer steps bump ``_version``) for nn.Parameters (or when cache=True: the caller
    keeps `w` alive and unchanged); never cached while a CUDA graph is being captured, so a captured training step
    re-packs inside the graph on every replay.  A cache entry belongs to ONE live tensor object (weak reference): the
    id / address / version of a freed parameter can all recur in the next model built by the same code, which must not
    get the old model's planes.  Weights changed behind autograd's back (``w.data.copy_()``, raw pointers) do not bump
    ``_version``: call ``invalidate_pack_cache()`` after such an edit."""
    is_param = isinstance(w, torch.nn.Parameter)
    want_wt = bool(w.requires_grad or need_wt)
    if is_param and cache is not False:
        if _record is not None:
            _record.append(("w", w, layout, want_wt))
        if _pinned:
            hit = _pinned_hit(w, ("w", layout, want_wt))
            if hit is not None:
                return hit
    capturing = torch.cuda.is_current_stream_capturing()
    if cache is None:
        cache = is_param
    if not cache or not _CACHE_OK:
        capturing = True            # same effect: neither look up nor store
    key = (w.data_ptr(), w._version, _epoch[0], tuple(w.shape), layout, want_wt)
    if not capturing:
        hit = _pack_cache.get(id(w))
        if hit is not None and hit.key == key and hit.owner is not None and hit.owner() is w:
            return hit
    wd = w.detach()
    if not wd.is_contiguous():
        wd = wd.contiguous()
    Cout, Cin, taps = _layout_dims(wd, layout)[:3]
    L = capi.lib()
    nbytes = int(L.sdf_spike_gemm_wq_bytes(Cout, Cin, taps))
    wq = torch.empty(nbytes, device=w.device, dtype=torch.int8)
    wscale = torch.empty(Cout, device=w.device, dtype=torch.float32)
    # transposed fp32 copy for the data-gradient GEMM, only when a backward pass can follow
    wt = torch.empty((Cin, taps * Cout), device=w.device, dtype=torch.float32) if want_wt else None
    _pack_launch(wd, layout, wq, wscale, wt)
    pw = PackedWeight(wq, wscale, wt, Cout, Cin, taps, key)
    if not capturing:
        wid = id(w)
        pw.owner = weakref.ref(w, lambda _r, wid=wid: _pack_cache.pop(wid, None) if (_pack_cache.get(wid) is not None
                                                                                     and _pack_cache[wid].owner is _r) else None)
        _pack_cache[wid] = pw
    return pw


_deconv_cache = {}


def _deconv_pack_launch(w, cin, packs):
    """(Re-)pack the four parity classes of a ConvTranspose2d weight into the buffers of `packs`."""
    import ctypes
    wd = w.detach()
    if cin != wd.shape[0]:
        wd = torch.nn.functional.pad(wd, (0, 0, 0, 0, 0, 0, 0, cin - wd.shape[0]))
    wd = wd.contiguous()
    Cin, Cout = wd.shape[0], wd.shape[1]
    L = capi.lib()
    for cls, pk in enumerate(packs):
        src, dh, dw = (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 4)()
        taps = int(L.sdf_spike_deconv_class_taps(cls, src, dh, dw))
        capi.call("sdf_spike_gemm_pack", capi.struct(
            "sdf_spike_gemm_pack_args", w=_ptr(wd), wq=_ptr(pk.wq), wscale=_ptr(pk.wscale), wt=None, wq_bytes=pk.wq.numel(), Cout=Cout,
            Cin=Cin, taps=taps, s_co=9, s_ci=Cout * 9, s_tap=1, tap_map=[int(src[i]) for i in range(taps)] + [0] * (9 - taps),
            stream=_stream()))


def pack_deconv_weight(w, cin=None, cache=None):
    """ConvTranspose2d weight (Cin, Cout, 3, 3) -> the four PackedWeights of sdf_spike_deconv_fwd (one per output parity
    class).  cin > w.shape[0]: zero input slices are appended first (decoder inputs concatenated up to a multiple of 16
    channels).  Cached per live parameter and version like pack_weight."""
    import ctypes
    cin = w.shape[0] if cin is None else cin
    is_param = isinstance(w, torch.nn.Parameter)
    if is_param and cache is not False:
        if _record is not None:
            _record.append(("deconv", w, cin))
        if _pinned:
            hit = _pinned_hit(w, ("deconv", cin))
            if hit is not None:
                return hit
    capturing = torch.cuda.is_current_stream_capturing() or not is_param or not _CACHE_OK or cache is False
    key = (w.data_ptr(), w._version, _epoch[0], tuple(w.shape), cin)
    if not capturing:
        hit = _deconv_cache.get(id(w))
        if hit is not None and hit[0] == key and hit[2]() is w:
            return hit[1]
    assert tuple(w.shape[2:]) == (3, 3)
    Cin, Cout = cin, w.shape[1]
    L = capi.lib()
    packs = []
    for cls in range(4):
        src, dh, dw = (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 4)()
        taps = int(L.sdf_spike_deconv_class_taps(cls, src, dh, dw))
        nbytes = int(L.sdf_spike_gemm_wq_bytes(Cout, Cin, taps))
        wq = torch.empty(nbytes, device=w.device, dtype=torch.int8)
        wscale = torch.empty(Cout, device=w.device, dtype=torch.float32)
        packs.append(PackedWeight(wq, wscale, None, Cout, Cin, taps, key))
    _deconv_pack_launch(w, cin, packs)
    if not capturing:
        wid = id(w)
        ref = weakref.ref(w, lambda _r, wid=wid: _deconv_cache.pop(wid, None) if (wid in _deconv_cache and _deconv_cache[wid][2] is _r) else None)
        _deconv_cache[wid] = (key, packs, ref)
    return packs


def spike_deconv_fwd(x_u8, packs, bias, want_stats=False, a_max=0):
    """ConvTranspose2d(k 3, stride 2, padding 1, output_padding 1) of 1-byte spikes (Nimg, H, W, Cin) -> fp32 (Nimg, 2H, 2W, Cout),
    BN partial sums [4 * N_PARTIAL, 2, Cout] when want_stats (one slab per parity class)."""
    Nimg, H, W, Cin = x_u8.shape
    Cout = packs[0].Cout
    assert x_u8.dtype == torch.uint8 and x_u8.is_contiguous() and packs[0].Cin == Cin
    out = torch.empty((Nimg, 2 * H, 2 * W, Cout), device=x_u8.device, dtype=torch.float32)
    part = torch.empty((4 * N_PARTIAL, 2, Cout), device=x_u8.device, dtype=torch.float32) if want_stats else None
    capi.call("sdf_spike_deconv_fwd", capi.struct(
        "sdf_spike_deconv_fwd_args", x=_ptr(x_u8), wq=[_ptr(p.wq) for p in packs], wscale=[_ptr(p.wscale) for p in packs],
        bias=_ptr(bias), out=_ptr(out), bn_partials=_ptr(part), n_partial_blocks=N_PARTIAL, Nimg=Nimg, H=H, W=W, Cin=Cin, Cout=Cout,
        a_max=a_max, stream=_stream()), algo_bytes=x_u8.numel() + 4 * out.numel())
    return out, part


def invalidate_pack_cache():
    """Drop every cached weight pack (after editing weights through .data / raw pointers)."""
    _pack_cache.clear()
    _deconv_cache.clear()
    bump_weights_epoch()


def spike_gemm_fwd(a_u8, pw, bias=None, want_stats=False, a_max=0):
    """a_u8 [rows, K] uint8 -> (out fp32 [rows, Cout], bn partials [N_PARTIAL, 2, Cout] or None).
    a_max: largest operand value (1 for spikes; 0 = any u8)."""
    rows, K = a_u8.shape
    assert a_u8.dtype == torch.uint8 and a_u8.is_contiguous() and K == pw.Cin and pw.taps == 1
    out = torch.empty((rows, pw.Cout), device=a_u8.device, dtype=torch.float32)
    part = torch.empty((N_PARTIAL, 2, pw.Cout), device=a_u8.device, dtype=torch.float32) if want_stats else None
    capi.call("sdf_spike_gemm_fwd", capi.struct(
        "sdf_spike_gemm_fwd_args", a=_ptr(a_u8), wq=_ptr(pw.wq), wscale=_ptr(pw.wscale), bias=_ptr(bias), out=_ptr(out),
        bn_partials=_ptr(part), n_partial_blocks=N_PARTIAL, rows=rows, K=K, Cout=pw.Cout, ld_out=pw.Cout, a_max=a_max, stream=_stream()),
        algo_bytes=rows * K + 4 * rows * pw.Cout)
    return out, part


def spike_conv_fwd(x_u8, pw, bias, kh, kw, stride, pad, want_stats=False, a_max=0):
    """x_u8 (Nimg, H, W, Cin) uint8 NHWC -> (out fp32 (Nimg, Ho, Wo, Cout), partials)."""
    Nimg, H, W, Cin = x_u8.shape
    assert x_u8.dtype == torch.uint8 and x_u8.is_contiguous() and Cin == pw.Cin and pw.taps == kh * kw
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    out = torch.empty((Nimg, Ho, Wo, pw.Cout), device=x_u8.device, dtype=torch.float32)
    part = torch.empty((N_PARTIAL, 2, pw.Cout), device=x_u8.device, dtype=torch.float32) if want_stats else None
    capi.call("sdf_spike_conv_fwd", capi.struct(
        "sdf_spike_conv_fwd_args", x=_ptr(x_u8), wq=_ptr(pw.wq), wscale=_ptr(pw.wscale), bias=_ptr(bias), out=_ptr(out),
        bn_partials=_ptr(part), n_partial_blocks=N_PARTIAL, Nimg=Nimg, H=H, W=W, Cin=Cin, Cout=pw.Cout, Ho=Ho, Wo=Wo,
        kh=kh, kw=kw, stride=stride, pad=pad, a_max=a_max, stream=_stream()),
        algo_bytes=x_u8.numel() + 4 * out.numel())
    return out, part


def gemm_tf32(a, b, bias=None):
    """a [rows, K] fp32 @ b [N, K]^T -> [rows, N] fp32, operands read as TF32."""
    rows, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and a.stride(1) == 1 and b.stride(1) == 1
    out = torch.empty((rows, N), device=a.device, dtype=torch.float32)
    capi.call("sdf_gemm_tf32", capi.struct(
        "sdf_gemm_tf32_args", a=_ptr(a), b=_ptr(b), bias=_ptr(bias), out=_ptr(out), rows=rows, K=K, N=N, lda=a.stride(0),
        ldb=b.stride(0), ld_out=N, stream=_stream()), algo_bytes=4 * (rows * K + rows * N))
    return out


_ws_cache = {}


def _workspace(nbytes, device):
    """One grow-only fp32 workspace per device for the split-row partial tiles (consumed within the same launch pair).
    Not shared across a CUDA-graph capture boundary: captured launches get their own buffer from the graph's pool."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(nbytes // 4, device=device, dtype=torch.float32)
    ws = _ws_cache.get(device)
    if ws is None or ws.numel() * 4 < nbytes:
        ws = torch.empty(nbytes // 4, device=device, dtype=torch.float32)
        _ws_cache[device] = ws
    return ws


def spike_wgrad(g, s_u8, out=None, s_max=0, want_db=False):
    """dW [Cout, K] = g[rows, Cout]^T @ s_u8[rows, K]; g fp32 (split into bf16 hi + lo operands), s uint8 (s_max = 1: the
    caller guarantees 0/1 spikes, which expand to bf16 with fewer instructions; 0: any byte value).
    want_db: also return the bias gradient g.sum(0) [Cout], accumulated by the same pass over g -> (dW, db)."""
    rows, Cout = g.shape
    K = s_u8.shape[1]
    assert s_u8.shape[0] == rows and s_u8.dtype == torch.uint8 and s_u8.is_contiguous() and g.stride(1) == 1
    nbytes = int(capi.lib().sdf_spike_wgrad_workspace_bytes(rows, Cout, K, 1))
    ws = _workspace(nbytes, g.device)
    acc = out is not None
    dw = out if acc else torch.empty((Cout, K), device=g.device, dtype=torch.float32)
    db = torch.empty(Cout, device=g.device, dtype=torch.float32) if want_db else None
    capi.call("sdf_spike_wgrad", capi.struct(
        "sdf_spike_wgrad_args", g=_ptr(g), s=_ptr(s_u8), dw=_ptr(dw), workspace=_ptr(ws), workspace_bytes=ws.numel() * 4,
        rows=rows, Cout=Cout, K=K, ldg=g.stride(0), accumulate=1 if acc else 0, s_max=s_max, stream=_stream(),
        db=_ptr(db)), algo_bytes=rows * (4 * Cout + K))
    return (dw, db) if want_db else dw


def spike_conv_wgrad(g, x_u8, kh, kw, stride, pad, s_max=0, want_db=False):
    """dW (Cout, Cin, kh, kw) of a convolution: g fp32 NHWC (Nimg, Ho, Wo, Cout), x_u8 NHWC (Nimg, H, W, Cin)."""
    Nimg, Ho, Wo, Cout = g.shape
    _, H, W, Cin = x_u8.shape
    assert g.is_contiguous() and x_u8.is_contiguous() and x_u8.dtype == torch.uint8
    pixels = Nimg * (-(-Ho // 2) * 2) * (-(-Wo // 16) * 16)       # whole 2 x 16 patches
    nbytes = int(capi.lib().sdf_spike_wgrad_workspace_bytes(pixels, Cout, Cin, kh * kw))
    ws = _workspace(nbytes, g.device)
    dw = torch.empty((Cout, Cin, kh, kw), device=g.device, dtype=torch.float32)
    db = torch.empty(Cout, device=g.device, dtype=torch.float32) if want_db else None
    capi.call("sdf_spike_conv_wgrad", capi.struct(
        "sdf_spike_conv_wgrad_args", g=_ptr(g), x=_ptr(x_u8), dw=_ptr(dw), workspace=_ptr(ws), workspace_bytes=ws.numel() * 4,
        Nimg=Nimg, H=H, W=W, Cin=Cin, Cout=Cout, Ho=Ho, Wo=Wo, kh=kh, kw=kw, stride=stride, pad=pad, accumulate=0,
        s_max=s_max, stream=_stream(), db=_ptr(db)), algo_bytes=4 * g.numel() + x_u8.numel())
    return (dw, db) if want_db else dw


def conv_dgrad_tf32(g, weight, H, W, pad, wd=None):
    """dX (Nimg, H, W, Cin) of a stride-1 convolution: g fp32 NHWC (Nimg, Ho, Wo, Cout), weight (Cout, Cin, kh, kw).
    wd: the [Cin][kh*kw*Cout] re-layout of the weight (PackedWeight.wt) if already at hand."""
    Nimg, Ho, Wo, Cout = g.shape
    _, Cin, kh, kw = weight.shape
    assert g.is_contiguous()
    if wd is None:
        wd = weight.detach().permute(1, 2, 3, 0).reshape(Cin, kh * kw * Cout).contiguous()
    out = torch.empty((Nimg, H, W, Cin), device=g.device, dtype=torch.float32)
    capi.call("sdf_conv_dgrad_tf32", capi.struct(
        "sdf_conv_dgrad_tf32_args", g=_ptr(g), wd=_ptr(wd), out=_ptr(out), Nimg=Nimg, H=H, W=W, Cin=Cin, Cout=Cout, Ho=Ho,
        Wo=Wo, kh=kh, kw=kw, pad=pad, stream=_stream()), algo_bytes=4 * (g.numel() + out.numel()))
    return out


def conv_dgrad_s2_tf32(g, weight, H, W, wd=None):
    """dX (Nimg, H, W, Cin) of a 3x3 / stride-2 / padding-1 convolution: g fp32 NHWC (Nimg, Ho, Wo, Cout), weight (Cout, Cin, 3, 3)
    (four parity-class launches of the TF32 implicit GEMM; wd as in conv_dgrad_tf32)."""
    Nimg, Ho, Wo, Cout = g.shape
    _, Cin, kh, kw = weight.shape
    assert g.is_contiguous() and (kh, kw) == (3, 3)
    if wd is None:
        wd = weight.detach().permute(1, 2, 3, 0).reshape(Cin, 9 * Cout).contiguous()
    out = torch.empty((Nimg, H, W, Cin), device=g.device, dtype=torch.float32)
    capi.call("sdf_conv_dgrad_s2_tf32", capi.struct(
        "sdf_conv_dgrad_s2_tf32_args", g=_ptr(g), wd=_ptr(wd), out=_ptr(out), Nimg=Nimg, H=H, W=W, Cin=Cin, Cout=Cout, Ho=Ho, Wo=Wo,
        stream=_stream()), algo_bytes=4 * (g.numel() + out.numel()))
    return out


def deconv_dgrad_weight(weight):
    """ConvTranspose2d weight (Cin, Cout, 3, 3) -> the B operand of deconv_dgrad_tf32: [Cin][9 * Cpad] with
    [ci][tap*Cpad + co] = W[ci, co, kh, kw], Cpad = Cout rounded up to 32 (whole 128-byte K chunks per tap), zero padded."""
    Cin, Cout = weight.shape[:2]
    Cpad = -(-Cout // 32) * 32
    wd = weight.detach().permute(0, 2, 3, 1)
    if Cpad != Cout:
        wd = torch.nn.functional.pad(wd, (0, Cpad - Cout))
    return wd.reshape(Cin, 9 * Cpad).contiguous()


def deconv_dgrad_tf32(g, weight, Cin=None, wd=None):
    """dX (Nimg, H, W, Cin) of ConvTranspose2d(3, stride 2, padding 1, output_padding 1): g fp32 NHWC (Nimg, 2H, 2W, Cout),
    weight (Cin_w, Cout, 3, 3); Cin > Cin_w: the extra (padding) channels of dX are zeros."""
    Nimg, Ho, Wo, Cout = g.shape
    Cin_w = weight.shape[0]
    Cin = Cin_w if Cin is None else Cin
    assert g.is_contiguous() and Ho % 2 == 0 and Wo % 2 == 0 and weight.shape[1] == Cout and tuple(weight.shape[2:]) == (3, 3)
    if wd is None:
        wd = deconv_dgrad_weight(weight)
    out = torch.empty((Nimg, Ho // 2, Wo // 2, Cin), device=g.device, dtype=torch.float32)
    capi.call("sdf_deconv_dgrad_tf32", capi.struct(
        "sdf_deconv_dgrad_tf32_args", g=_ptr(g), wd=_ptr(wd), out=_ptr(out), Nimg=Nimg, H=Ho // 2, W=Wo // 2, Cin=Cin, Cin_w=Cin_w,
        Cout=Cout, stream=_stream()), algo_bytes=4 * (g.numel() + out.numel()))
    return out


def spike_deconv_wgrad(g, x_u8, Cin_w=None, s_max=0, want_db=False):
    """dW (Cin_w, Cout, 3, 3) of ConvTranspose2d(3, stride 2, padding 1, output_padding 1): g fp32 NHWC (Nimg, 2H, 2W, Cout),
    x_u8 NHWC (Nimg, H, W, Cin >= Cin_w) spikes.  want_db: also the bias gradient g.sum((0, 1, 2))."""
    Nimg, Ho, Wo, Cout = g.shape
    _, H, W, Cin = x_u8.shape
    Cin_w = Cin if Cin_w is None else Cin_w
    assert g.is_contiguous() and x_u8.is_contiguous() and x_u8.dtype == torch.uint8 and (Ho, Wo) == (2 * H, 2 * W)
    nbytes = int(capi.lib().sdf_spike_deconv_wgrad_workspace_bytes(Nimg, H, W, Cout, Cin))
    ws = _workspace(nbytes, g.device)
    dw = torch.empty((Cin_w, Cout, 3, 3), device=g.device, dtype=torch.float32)
    db = torch.empty(Cout, device=g.device, dtype=torch.float32) if want_db else None
    capi.call("sdf_spike_deconv_wgrad", capi.struct(
        "sdf_spike_deconv_wgrad_args", g=_ptr(g), x=_ptr(x_u8), dw=_ptr(dw), workspace=_ptr(ws), workspace_bytes=ws.numel() * 4,
        Nimg=Nimg, H=H, W=W, Cin=Cin, Cin_w=Cin_w, Cout=Cout, s_max=s_max, stream=_stream(), db=_ptr(db)),
        algo_bytes=4 * g.numel() + x_u8.numel())
    return (dw, db) if want_db else dw
