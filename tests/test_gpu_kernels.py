"""GPU parity tests, kernel level: every C-ABI kernel against the CPU oracle (oracle/port.py) on
identical seeded inputs.  Bars (BASELINE.json north_star): spikes and index maps bit-exact given
identical inputs; membrane potentials <= 1e-4 relative; spike-flip rate <= 1e-4 where a library
GEMM / BN reduction order sits between input and spike."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import port
from helpers import flip_rate, rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from sdformerflow_b200 import ops, capi
    return ops, capi


def _cfg(ops, capi, kind="lif", v_th=0.5, v_reset=None, tau=2.0, detach=True):
    k = {"lif": capi.SDF_NEURON_LIF, "if": capi.SDF_NEURON_IF, "plif": capi.SDF_NEURON_PLIF}[kind]
    return ops.NeuronCfg(kind=k, v_th=v_th, v_reset=v_reset, tau=tau, detach_reset=detach)


def _spec(T, kind="lif", v_th=0.5, v_reset=None, tau=2.0, detach=True):
    return port.NeuronSpec(T, kind, v_th, v_reset, tau, detach)


# ---------------------------------------------------------------------------------------------
# K1 / K2
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T", [1, 2, 3, 4, 5, 10, 20])
@pytest.mark.parametrize("kind,v_reset,tau", [("lif", None, 2.0), ("lif", 0.0, 2.0), ("lif", -0.2, 3.0), ("if", None, 2.0)])
def test_lif_fwd_bit_exact(T, kind, v_reset, tau):
    ops, capi = _ops()
    g = torch.Generator().manual_seed(T * 7 + 1)
    x = torch.randn(T, 3, 5, 8, 12, generator=g) * 0.5 + 0.15
    rec = []
    ref = port.lif_multistep(x, _spec(T, kind, 0.5, v_reset, tau), kind, None, rec)
    s, h = ops.neuron_debug(x.to(DEV), _cfg(ops, capi, kind, 0.5, v_reset, tau))
    assert torch.equal(s.cpu(), ref)
    assert torch.equal(h.cpu(), rec[0])  # same fp32 op order, FMA contraction off => bit-exact membrane


def test_lif_fwd_odd_sizes_and_dtypes():
    ops, capi = _ops()
    g = torch.Generator().manual_seed(3)
    for shape in [(10, 7), (10, 1, 333), (4, 2, 3, 5)]:           # not multiples of 4 -> scalar path
        x = torch.randn(*shape, generator=g) * 0.6
        ref = port.lif_multistep(x, _spec(shape[0]), "lif")
        s, _ = ops.neuron_debug(x.to(DEV), _cfg(ops, capi))
        assert torch.equal(s.cpu(), ref)
    x = torch.randn(10, 4, 96, generator=g) * 0.6
    ref = port.lif_multistep(x, _spec(10), "lif")
    for dt, tdt in ((capi.SDF_SPIKE_U8, torch.uint8), (capi.SDF_SPIKE_BF16, torch.bfloat16)):
        s, _ = ops.neuron_debug(x.to(DEV), _cfg(ops, capi), spike_dtype=dt)
        assert s.dtype == tdt
        assert torch.equal(s.float().cpu(), ref)


def test_lif_time_strided_layout_matches_permute():
    """(B, D, H, W, C) with time = D (time_dim=1) == reference's x.permute(1,0,2,3,4) call (:845)."""
    ops, capi = _ops()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 10, 4, 6, 32, generator=g) * 0.5
    ref = port.lif_multistep(x.permute(1, 0, 2, 3, 4), _spec(10, v_th=0.1), "lif").permute(1, 0, 2, 3, 4)
    s, _ = ops.neuron_debug(x.to(DEV), _cfg(ops, capi, v_th=0.1), time_dim=1)
    assert torch.equal(s.cpu(), ref)


@pytest.mark.parametrize("T", [2, 4, 5, 10, 7, 20])
@pytest.mark.parametrize("kind,v_reset,detach", [("lif", None, True), ("lif", None, False), ("lif", 0.0, True),
                                                  ("lif", 0.1, False), ("if", None, True)])
def test_lif_bwd_matches_autograd(T, kind, v_reset, detach):
    ops, capi = _ops()
    g = torch.Generator().manual_seed(T + 11)
    x = (torch.randn(T, 2, 6, 40, generator=g) * 0.5 + 0.1).requires_grad_(True)
    go = torch.randn(T, 2, 6, 40, generator=g)
    port.lif_multistep(x, _spec(T, kind, 0.5, v_reset, 2.0, detach), kind).backward(go)
    xg = x.detach().to(DEV).requires_grad_(True)
    ops.neuron(xg, _cfg(ops, capi, kind, 0.5, v_reset, 2.0, detach)).backward(go.to(DEV))
    assert torch.allclose(xg.grad.cpu(), x.grad, rtol=1e-5, atol=1e-6)


def test_plif_fwd_bwd():
    ops, capi = _ops()
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(10, 3, 64, generator=g) * 0.5 + 0.1).requires_grad_(True)
    w = torch.tensor(0.3, requires_grad=True)
    go = torch.randn(10, 3, 64, generator=g)
    ref = port.lif_multistep(x, _spec(10, "plif"), "plif", w)
    ref.backward(go)
    xg = x.detach().to(DEV).requires_grad_(True)
    wg = w.detach().to(DEV).requires_grad_(True)
    out = ops.neuron(xg, _cfg(ops, capi, "plif"), plif_w=wg)
    out.backward(go.to(DEV))
    assert flip_rate(out.cpu(), ref.detach()) <= 1e-3   # k = sigmoid(w) rounds once on each side
    assert torch.allclose(xg.grad.cpu(), x.grad, rtol=1e-3, atol=1e-4)
    assert abs(wg.grad.item() - w.grad.item()) <= 2e-3 * max(1.0, abs(w.grad.item()))


@pytest.mark.parametrize("T", [2, 5, 10, 6])
def test_psn_fwd_bwd(T):
    ops, capi = _ops()
    g = torch.Generator().manual_seed(T)
    x = (torch.randn(T, 4, 9, 32, generator=g) * 0.7).requires_grad_(True)
    W = (torch.eye(T) * 0.8 + torch.randn(T, T, generator=g) * 0.2).requires_grad_(True)
    b = (torch.full((T, 1), -0.1) + torch.randn(T, 1, generator=g) * 0.05).requires_grad_(True)
    go = torch.randn(T, 4, 9, 32, generator=g)
    rec = []
    ref = port.psn_forward(x, W, b, 2.0, rec)
    ref.backward(go)
    xg, Wg, bg = (t.detach().to(DEV).requires_grad_(True) for t in (x, W, b))
    cfg = ops.NeuronCfg(kind=capi.SDF_NEURON_IF, v_th=0.0, v_reset=None)
    out = ops.psn(xg, Wg, bg, cfg)
    out.backward(go.to(DEV))
    # sgemm summation order differs from the in-register FMA chain: allow threshold ties to flip
    assert flip_rate(out.cpu(), ref.detach()) <= 1e-4
    assert torch.allclose(xg.grad.cpu(), x.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(Wg.grad.cpu(), W.grad, rtol=1e-3, atol=1e-3)
    assert torch.allclose(bg.grad.cpu(), b.grad, rtol=1e-3, atol=1e-3)


# ---------------------------------------------------------------------------------------------
# K6 BatchNorm (+ fused neuron)
# ---------------------------------------------------------------------------------------------
def _bn(C, seed, train):
    g = torch.Generator().manual_seed(seed)
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    bn.train(train)
    return bn


@pytest.mark.parametrize("C,T", [(96, 10), (384, 10), (3072, 10), (96, 20), (192, 5)])
@pytest.mark.parametrize("train", [False, True])
def test_bn_neuron_fwd_bwd(C, T, train):
    """neuron(BN(u)) on channels-last rows vs sn(bn(u.permute(0,1,4,2,3)).permute(0,1,3,4,2)); T = 20 is the 20-bin input of
    the reference's MDR configs (vector path of K2 with the BN partial sums)."""
    ops, capi = _ops()
    import copy
    g = torch.Generator().manual_seed(C)
    B, H, W = 2, 3, 5
    u = (torch.randn(T, B, H, W, C, generator=g) * 0.8 + 0.2).requires_grad_(True)
    go = torch.randn(T, B, H, W, C, generator=g)
    bn_ref = _bn(C, 1, train)
    bn_gpu = copy.deepcopy(bn_ref).to(DEV)
    y = bn_ref(u.permute(0, 1, 4, 2, 3).flatten(0, 1)).view(T, B, C, H, W).permute(0, 1, 3, 4, 2)
    rec = []
    ref = port.lif_multistep(y, _spec(T, v_th=0.3), "lif", None, rec)
    ref.backward(go)
    ug = u.detach().to(DEV).requires_grad_(True)
    out = ops.bn_neuron(ug, bn_gpu, _cfg(ops, capi, v_th=0.3), 0)
    out.backward(go.to(DEV))
    assert flip_rate(out.cpu(), ref.detach()) <= 1e-4
    if train:
        assert torch.allclose(bn_gpu.running_mean.cpu(), bn_ref.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(bn_gpu.running_var.cpu(), bn_ref.running_var, rtol=1e-5, atol=1e-6)
        assert int(bn_gpu.num_batches_tracked) == int(bn_ref.num_batches_tracked)
    scale = x_scale = u.grad.abs().max()
    assert (ug.grad.cpu() - u.grad).abs().max() <= 2e-3 * x_scale  # a flipped spike near threshold moves a few grads
    assert torch.allclose(bn_gpu.weight.grad.cpu(), bn_ref.weight.grad, rtol=2e-3, atol=2e-3 * bn_ref.weight.grad.abs().max())
    assert torch.allclose(bn_gpu.bias.grad.cpu(), bn_ref.bias.grad, rtol=2e-3, atol=2e-3 * bn_ref.bias.grad.abs().max())
    del scale


@pytest.mark.parametrize("train", [False, True])
def test_bn_residual_fwd_bwd(train):
    ops, capi = _ops()
    import copy
    g = torch.Generator().manual_seed(9)
    u = torch.randn(2, 10, 5, 4, 192, generator=g).requires_grad_(True)
    r = torch.randn(2, 10, 5, 4, 192, generator=g).requires_grad_(True)
    go = torch.randn(2, 10, 5, 4, 192, generator=g)
    bn_ref = _bn(192, 2, train)
    bn_gpu = copy.deepcopy(bn_ref).to(DEV)
    ref = bn_ref(u.permute(0, 4, 1, 2, 3).reshape(2, 192, 50, 4)).view(2, 192, 10, 5, 4).permute(0, 2, 3, 4, 1) + r
    ref.backward(go)
    ug, rg = (t.detach().to(DEV).requires_grad_(True) for t in (u, r))
    out = ops.bn_residual(ug, bn_gpu, rg)
    out.backward(go.to(DEV))
    assert rel_err(out.cpu(), ref.detach()) <= 1e-5
    assert rel_err(ug.grad.cpu(), u.grad) <= 1e-4
    assert torch.equal(rg.grad.cpu(), go)
    assert rel_err(bn_gpu.weight.grad.cpu(), bn_ref.weight.grad) <= 1e-4
    assert rel_err(bn_gpu.bias.grad.cpu(), bn_ref.bias.grad) <= 1e-4


# ---------------------------------------------------------------------------------------------
# window index algebra (bit-exact)
# ---------------------------------------------------------------------------------------------
GEOMS = [
    # B, D, H, W, window, shift
    (2, 10, 24, 32, (2, 9, 9), (0, 0, 0)),
    (2, 10, 24, 32, (2, 9, 9), (1, 4, 4)),
    (1, 5, 16, 16, (2, 8, 8), (1, 4, 4)),      # D padded 5 -> 6 (MDR config)
    (1, 5, 8, 8, (2, 8, 8), (1, 4, 4)),        # H,W == window -> shift clamped to (1,0,0)
    (3, 4, 7, 10, (2, 3, 4), (1, 1, 2)),
    (1, 10, 9, 18, (2, 9, 9), (1, 4, 4)),      # shift_h clamped to 0 (288x384 stage 4)
    (2, 8, 13, 11, (4, 12, 12), (2, 6, 6)),
]


def _ref_windows(x, window, shift):
    B, D, H, W, C = x.shape
    ws, ss = port.get_window_size((D, H, W), window, shift)
    pd, pb, pr = (ws[0] - D % ws[0]) % ws[0], (ws[1] - H % ws[1]) % ws[1], (ws[2] - W % ws[2]) % ws[2]
    xp = F.pad(x, (0, 0, 0, pr, 0, pb, 0, pd))
    if any(s > 0 for s in ss):
        xp = torch.roll(xp, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
    return port.window_partition_v2(xp, ws), ws, ss, xp.shape[1:4]


@pytest.mark.parametrize("geom", GEOMS)
def test_window_gather_scatter_bit_exact(geom):
    ops, capi = _ops()
    B, D, H, W, window, shift = geom
    C = 8
    x = torch.arange(B * D * H * W * C, dtype=torch.float32).view(B, D, H, W, C) + 1.0
    xw_ref, ws, ss, (Dp, Hp, Wp) = _ref_windows(x, window, shift)
    g = ops.WindowGeom.get(B, D, H, W, window, shift, DEV)
    assert g.window == ws and g.shift == ss
    xw = ops.window_gather(x.to(DEV), g)
    assert xw.shape == xw_ref.shape
    assert torch.equal(xw.cpu(), xw_ref)
    # reverse: view -> window_reverse -> roll back -> crop (:810-820), plus residual
    y = torch.randn(g.rows, C)
    yr = port.window_reverse(y.view(-1, *(ws + (C,))), ws, B, Dp, Hp, Wp)
    if any(s > 0 for s in ss):
        yr = torch.roll(yr, shifts=ss, dims=(1, 2, 3))
    yr = yr[:, :D, :H, :W, :] + x
    out = ops.window_scatter(y.to(DEV), g, res=x.to(DEV))
    assert torch.equal(out.cpu(), yr)
    # region ids reproduce compute_mask
    if any(s > 0 for s in ss):
        mask = port.compute_mask(Dp, Hp, Wp, ws, ss)
        reg = g.region.cpu().view(g.nW, g.N).float()
        mine = (reg.unsqueeze(1) - reg.unsqueeze(2) != 0).float() * -100.0
        assert torch.equal(mine, mask)


@pytest.mark.parametrize("geom", GEOMS[:5])
def test_lif_window_fwd_bwd(geom):
    ops, capi = _ops()
    B, D, H, W, window, shift = geom
    C = 32
    g0 = torch.Generator().manual_seed(2)
    x = (torch.randn(B, D, H, W, C, generator=g0) * 0.6 + 0.1).requires_grad_(True)
    xw_ref, ws, ss, _ = _ref_windows(x, window, shift)
    rec = []
    ref = port.lif_multistep(xw_ref, _spec(ws[0], v_th=0.2), "lif", None, rec)
    go = torch.randn(ref.shape, generator=g0)
    ref.backward(go)
    g = ops.WindowGeom.get(B, D, H, W, window, shift, DEV)
    s, h = ops.lif_window_debug(x.detach().to(DEV), g, _cfg(ops, capi, v_th=0.2))
    assert torch.equal(s.cpu(), ref.detach())
    assert torch.equal(h.cpu(), rec[0].detach())
    xg = x.detach().to(DEV).requires_grad_(True)
    ops.lif_window(xg, g, _cfg(ops, capi, v_th=0.2)).backward(go.to(DEV))
    assert torch.allclose(xg.grad.cpu(), x.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("H,W", [(8, 12), (7, 9)])
@pytest.mark.parametrize("apply_neuron", [True, False])
def test_lif_merge_fwd_bwd(H, W, apply_neuron):
    ops, capi = _ops()
    g0 = torch.Generator().manual_seed(4)
    B, D, C = 2, 10, 32
    x = (torch.randn(B, D, H, W, C, generator=g0) * 0.6 + 0.1).requires_grad_(True)
    xp = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    cat = torch.cat([xp[:, :, 0::2, 0::2], xp[:, :, 1::2, 0::2], xp[:, :, 0::2, 1::2], xp[:, :, 1::2, 1::2]], -1)
    ref = port.lif_multistep(cat.permute(1, 0, 2, 3, 4), _spec(D, v_th=0.2), "lif").permute(1, 0, 2, 3, 4) if apply_neuron else cat
    go = torch.randn(ref.shape, generator=g0)
    ref.backward(go)
    xg = x.detach().to(DEV).requires_grad_(True)
    out = ops.lif_merge(xg, _cfg(ops, capi, v_th=0.2), apply_neuron)
    out.backward(go.to(DEV))
    assert torch.equal(out.detach().cpu(), ref.detach())
    assert torch.allclose(xg.grad.cpu(), x.grad, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# K5 QK-gate
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wd,wh,ww,nH,M", [(2, 9, 9, 3, 8), (2, 3, 4, 6, 10), (4, 3, 3, 3, 4), (2, 8, 8, 24, 2)])
def test_qkgate_core_bit_exact(wd, wh, ww, nH, M):
    """Given identical q_pre/k_pre and folded BN vectors, the gate spikes, the three membranes and the
    permuted output order are bit-exact w.r.t. the reference op sequence (:672-710)."""
    ops, capi = _ops()
    C, P = nH * 32, wh * ww
    g0 = torch.Generator().manual_seed(wd * 100 + nH)
    q_pre = torch.randn(wd, M, wh, ww, C, generator=g0)
    k_pre = torch.randn(wd, M, wh, ww, C, generator=g0)
    qs, ks = torch.rand(C, generator=g0) + 0.5, torch.rand(C, generator=g0) + 0.5
    qh, kh = torch.randn(C, generator=g0) * 0.2, torch.randn(C, generator=g0) * 0.2
    pos = torch.randn(1, nH, wd * P, 32, generator=g0) * 0.2
    spec = _spec(wd, v_th=0.3)
    # oracle: same op order as the reference, BN pre-folded as x*scale + shift with an fma (torch.addcmul)
    rec = []
    q = torch.addcmul(qh, q_pre, qs)
    k = torch.addcmul(kh, k_pre, ks) + pos.reshape(wd, 1, wh, ww, C)
    q = port.lif_multistep(q, spec, "lif", None, rec)
    k = port.lif_multistep(k, spec, "lif", None, rec)
    qr, kr = q.reshape(wd, M, nH, -1, 32), k.reshape(M, nH, -1, 32)
    att = port.lif_multistep(qr.sum(dim=-1, keepdim=True), spec, "lif", None, rec)
    attn = kr.mul(att.reshape(M, nH, -1, 1))
    ref = attn.reshape(M, nH, wd, wh, ww, 32).permute(2, 0, 3, 4, 1, 5).reshape(wd, M, wh, ww, C)
    rows = wd * M * P
    gate, q_h, k_h, a_h = ops.qkgate_debug(q_pre.view(rows, C).to(DEV), k_pre.view(rows, C).to(DEV), qs.to(DEV), qh.to(DEV),
                                           ks.to(DEV), kh.to(DEV), pos.to(DEV), _cfg(ops, capi, v_th=0.3), wd, M, P, nH)
    # torch.addcmul on CPU may or may not fuse; accept either exact equality or a vanishing flip rate
    assert flip_rate(gate.cpu().view_as(ref), ref) <= 1e-5
    assert rel_err(q_h.cpu().view_as(rec[0]), rec[0]) <= 1e-6
    assert rel_err(k_h.cpu().view_as(rec[1]), rec[1]) <= 1e-6
    assert rel_err(a_h.cpu().view(-1), rec[2].reshape(-1)) <= 1e-6


@pytest.mark.parametrize("train", [False, True])
@pytest.mark.parametrize("wd,wh,ww,nH,M", [(2, 3, 4, 3, 6), (2, 9, 9, 3, 4)])
def test_qk_attention_module_fwd_bwd(train, wd, wh, ww, nH, M):
    """Spiking_QK_WindowAttention3D end to end (proj_sn, q/k GEMMs, BN, gate, proj, proj_bn) vs the port."""
    from sdformerflow_b200.STSwinNet_SNN import Spiking_swin_transformer3D as prod
    from oracle import synth
    C = nH * 32
    kw = {"num_steps": 10, "v_reset": None, "v_th": 0.2, "neuron_type": "lif", "surrogate_fun": "surrogate.ATan()",
          "tau": 2.0, "detach_reset": True, "spike_norm": "BN"}
    m = prod.Spiking_QK_WindowAttention3D(C, (wd, wh, ww), (0, 0, 0), nH, norm="BN", **kw)
    sd = synth.synth_state_dict(m.state_dict(), seed=11)
    m.load_state_dict(sd)
    m.train(train).to(DEV)
    P = port.params_from_state_dict({"a." + k: v for k, v in sd.items()}, requires_grad=True)
    g0 = torch.Generator().manual_seed(1)
    x = (torch.randn(wd, M, wh, ww, C, generator=g0) * 0.7 + 0.1).requires_grad_(True)
    ref, _ = port.qk_window_attention(x, P, "a", nH, port.NeuronSpec(10, "lif", 0.2, None, 2.0, True), port.BNMode(train))
    go = torch.randn(ref.shape, generator=g0)
    ref.backward(go)
    xg = x.detach().to(DEV).requires_grad_(True)
    out, _ = m(xg)
    out.backward(go.to(DEV))
    # proj_bn output is continuous: spikes upstream may flip at ties (GEMM order), so compare in norm
    err = (out.detach().cpu() - ref.detach()).abs()
    assert (err > 1e-4 * ref.abs().max()).float().mean().item() <= 2e-3
    gerr = (xg.grad.cpu() - x.grad).abs()
    assert (gerr > 1e-3 * x.grad.abs().max()).float().mean().item() <= 5e-3
    for name in ("linear_q.weight", "proj.weight", "positional_encoding", "bn_k.norm_layer.weight"):
        a, b = dict(m.named_parameters())[name].grad.cpu(), P["a." + name].grad
        assert ((a - b).abs() > 2e-2 * b.abs().max()).float().mean().item() <= 1e-2, name


# ---------------------------------------------------------------------------------------------
# fp32-faithful spike GEMM / conv on tensor cores (TF32 x 2 weight split)
# ---------------------------------------------------------------------------------------------
def test_split_tf32_is_exactly_representable_and_tight():
    ops, capi = _ops()
    w = torch.randn(4096, generator=torch.Generator().manual_seed(0)).to(DEV) * 0.1
    hi, lo = ops.split_tf32(w)
    for t in (hi, lo):
        assert int((t.view(torch.int32) & 0x1FFF).abs().max()) == 0          # 13 low mantissa bits clear
    assert ((w - hi - lo).abs() <= w.abs() * 2.0 ** -21).all()


@pytest.mark.parametrize("mode", ["tf32x2", "fp32"])
def test_spike_linear_and_conv_are_fp32_grade(mode):
    ops, capi = _ops()
    old = ops.GEMM_MODE
    ops.GEMM_MODE = mode
    try:
        g = torch.Generator().manual_seed(1)
        s = (torch.rand(4000, 384, generator=g) < 0.35).float().to(DEV)
        W = (torch.randn(1536, 384, generator=g) * 0.07).to(DEV)
        b = torch.randn(1536, generator=g).to(DEV)
        ref = (s.double() @ W.double().t() + b.double())
        y = ops.spike_linear(s, W, b)
        assert ((y.double() - ref).abs().max() / ref.abs().max()).item() <= 2e-6
        x = (torch.rand(6, 96, 36, 48, generator=g) < 0.3).float().to(DEV)
        Wc = (torch.randn(96, 96, 3, 3, generator=g) * 0.05).to(DEV)
        refc = F.conv2d(x.double(), Wc.double(), None, stride=2, padding=1)
        yc = ops.spike_conv2d(x, Wc, None, 2, 1)
        assert ((yc.double() - refc).abs().max() / refc.abs().max()).item() <= 2e-6
        Wt = (torch.randn(96, 48, 3, 3, generator=g) * 0.05).to(DEV)
        reft = F.conv_transpose2d(x.double(), Wt.double(), None, stride=2, padding=1, output_padding=1)
        yt = ops.spike_conv2d(x, Wt, None, 2, 1, True, 1)
        assert ((yt.double() - reft).abs().max() / reft.abs().max()).item() <= 2e-6
        # backward (single-pass TF32): gradients within 2e-3 of fp64
        s.requires_grad_(True)
        Wp = W.clone().requires_grad_(True)
        go = torch.randn(4000, 1536, generator=g).to(DEV)
        ops.spike_linear(s, Wp, b).backward(go)
        gs_ref = go.double() @ W.double()
        gw_ref = go.double().t() @ s.detach().double()
        tol = 3e-3 if mode == "tf32x2" else 1e-5
        assert ((s.grad.double() - gs_ref).abs().max() / gs_ref.abs().max()).item() <= tol
        assert ((Wp.grad.double() - gw_ref).abs().max() / gw_ref.abs().max()).item() <= tol
    finally:
        ops.GEMM_MODE = old


@pytest.mark.parametrize("Cin,Cout,H,W", [(2, 48, 17, 23), (1, 8, 5, 4), (4, 64, 9, 16), (2, 48, 96, 128), (3, 128, 7, 9)])
def test_small_cin_conv_fwd_bwd(Cin, Cout, H, W):
    """Direct 3x3 conv of the patch-embed head vs F.conv2d (fp64 reference), and its gradients."""
    ops, capi = _ops()
    g = torch.Generator().manual_seed(Cin * 10 + Cout)
    x = (torch.rand(6, H, W, Cin, generator=g) * (torch.rand(6, H, W, Cin, generator=g) < 0.3)).requires_grad_(True)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.3).requires_grad_(True)
    b = torch.randn(Cout, generator=g).requires_grad_(True)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), b.double(), stride=1, padding=1).permute(0, 2, 3, 1)
    go = torch.randn(ref.shape, generator=g)
    ref.backward(go.double())
    xg, wg, bg = (t.detach().to(DEV).requires_grad_(True) for t in (x, w, b))
    y = ops.conv3x3_small_cin(xg, wg, bg)
    y.backward(go.to(DEV))
    assert ((y.detach().cpu().double() - ref.detach()).abs().max() / ref.abs().max()).item() <= 1e-6
    assert ((wg.grad.cpu().double() - w.grad.double()).abs().max() / w.grad.abs().max()).item() <= 3e-3
    assert ((xg.grad.cpu().double() - x.grad.double()).abs().max() / x.grad.abs().max()).item() <= 3e-3
    assert ((bg.grad.cpu().double() - b.grad.double()).abs().max() / b.grad.abs().max()).item() <= 1e-4
    # the model's case: the input needs no gradient -> dW / db from the library's own direct fp32 kernel, deterministic
    grads = []
    for _ in range(2):
        w2, b2 = (t.detach().to(DEV).requires_grad_(True) for t in (w, b))
        ops.conv3x3_small_cin(x.detach().to(DEV), w2, b2).backward(go.to(DEV))
        grads.append((w2.grad.clone(), b2.grad.clone()))
    assert ((grads[0][0].cpu().double() - w.grad.double()).abs().max() / w.grad.abs().max()).item() <= 1e-5
    assert ((grads[0][1].cpu().double() - b.grad.double()).abs().max() / b.grad.abs().max()).item() <= 1e-5
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])


# ---------------------------------------------------------------------------------------------
# input pipeline: polarity split + min-max normalisation + bins->steps regroup
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,bins,steps,H,W", [(2, 10, 10, 24, 40), (1, 10, 5, 17, 23), (3, 5, 5, 16, 16), (1, 20, 10, 8, 12)])
def test_voxel_prepare_matches_script_lines(B, bins, steps, H, W):
    """ops.prepare_voxels == train_flow_parallel_supervised_SNN.py:261-265,278-284 followed by the patch embedding's regroup
    (Spiking_modules.py:1772-1786), bit for bit."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(B * 100 + bins)
    raw = torch.rand(B, bins, H, W, generator=g) * 3.0 * (torch.rand(B, bins, H, W, generator=g) < 0.3) \
        * (torch.randint(0, 2, (B, bins, H, W), generator=g) * 2 - 1)
    # reference script lines (CPU torch)
    chunk = raw.clone()
    neg = F.relu(-chunk)
    pos = F.relu(chunk)
    chunk = torch.cat((pos.unsqueeze(2), neg.unsqueeze(2)), dim=2)
    mn, mx = torch.min(chunk[chunk != 0]), torch.max(chunk[chunk != 0])
    if not mn == mx:
        chunk[chunk != 0] = (chunk[chunk != 0] - mn) / (mx - mn)
    ref = port.regroup_events(chunk, bins, steps)                     # (steps, B, num_ch, H, W)
    got = ops.prepare_voxels(raw.to(DEV), steps)                      # (B, steps, H, W, num_ch)
    assert torch.equal(got.cpu(), ref.permute(1, 0, 3, 4, 2).contiguous())
    # the model-side regroup alone, on what the scripts hand to the model
    got2 = ops.regroup_voxels(chunk.to(DEV), steps)
    assert torch.equal(got2.cpu(), ref.permute(1, 0, 3, 4, 2).contiguous())
    # all-equal non-zero values: min == max, the reference skips the normalisation
    flat = (torch.rand(1, bins, H, W, generator=g) < 0.2).float() * 0.7
    out = ops.prepare_voxels(flat.to(DEV), steps).cpu()
    assert int((out != 0).sum()) == int((flat != 0).sum()) and bool(((out == 0) | (out == flat.max())).all())
