"""GPU parity tests of K3/K4 (tcgen05 Q K^T V attention): integer Q K^T counts bit-exact, attn and output
against the reference formula (Spiking_swin_transformer3D.py:320-363), gradients against autograd, and the
SEW / SDSA Swin stage against the fixtures generated from the unmodified reference."""
import pytest
import torch

from oracle import port, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # wd, wh, ww, nH, B, nWin/sample, masked
    (2, 3, 4, 2, 2, 3, True),
    (2, 9, 9, 3, 2, 4, True),
    (2, 9, 9, 3, 1, 5, False),
    (2, 8, 8, 6, 1, 2, True),
    (2, 10, 10, 3, 1, 2, False),     # N = 200: two key tiles, second one partial
    (4, 12, 12, 3, 1, 2, True),      # N = 576: 5 M-tiles x 3 key tiles
]


def _inputs(wd, wh, ww, nH, B, nW, masked, seed=0):
    g = torch.Generator().manual_seed(seed + wd * wh)
    N, P, C = wd * wh * ww, wh * ww, nH * 32
    M = B * nW
    q, k, v = ((torch.rand(wd, M, P, C, generator=g) < r).to(torch.uint8) for r in (0.25, 0.3, 0.35))
    table = torch.randn((2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), nH, generator=g) * 0.2
    region = torch.randint(0, 3, (nW, N), generator=g).to(torch.uint8) if masked else None
    return q, k, v, table, region, M, N, P, C


def _reference(q, k, v, table, region, M, N, nH, window, scale):
    """(q*scale) @ k^T + bias + mask, then @ v and the reshape/permute of :362-363, in float64-free fp32."""
    wd, wh, ww = window
    C = nH * 32
    qf, kf, vf = (t.float().reshape(M, nH, N, 32) for t in (q, k, v))
    S = (qf @ kf.transpose(-2, -1))
    idx = port.relative_position_index(window)[:N, :N].reshape(-1)
    bias = table[idx].reshape(N, N, -1).permute(2, 0, 1).contiguous()
    attn = (qf * scale) @ kf.transpose(-2, -1) + bias.unsqueeze(0)
    if region is not None:
        nW = region.shape[0]
        r = region.float()
        mask = (r.unsqueeze(1) != r.unsqueeze(2)).float() * -100.0
        attn = (attn.view(M // nW, nW, nH, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, nH, N, N)
    x = (attn @ vf).reshape(M, nH, wd, wh, ww, 32).permute(2, 0, 3, 4, 1, 5).reshape(wd * M * wh * ww, C)
    return S, attn, x


@pytest.mark.parametrize("case", CASES)
def test_qktv_forward(case):
    from sdformerflow_b200 import ops
    wd, wh, ww, nH, B, nW, masked = case
    q, k, v, table, region, M, N, P, C = _inputs(*case)
    S, attn, x = _reference(q, k, v, table, region, M, N, nH, (wd, wh, ww), 0.125)
    out, s_dbg, a_dbg = ops.qktv_debug(q.to(DEV), k.to(DEV), v.to(DEV), table.to(DEV),
                                       None if region is None else region.to(DEV), M, nH, nW, (wd, wh, ww), 0.125)
    assert torch.equal(s_dbg.cpu().view(M, nH, N, N), S.to(torch.int32))          # binary spike products: bit-exact
    assert torch.allclose(a_dbg.cpu().view(M, nH, N, N), attn, rtol=0, atol=1e-5)
    err = (out.cpu() - x).abs().max() / x.abs().max()
    assert err.item() <= 2e-6, err.item()


# window sizes that cover every key-padding template of the small-window kernel (N <= 176: bias resident in TMEM),
# one and two M-tiles, full and partial key tiles
FAST_CASES = [
    # wd, wh, ww, nH, B, nWin/sample
    (1, 4, 4, 2, 2, 3),      # N = 16
    (2, 3, 4, 2, 2, 3),      # N = 24  -> 32
    (3, 4, 4, 1, 1, 5),      # N = 48
    (2, 5, 5, 3, 1, 4),      # N = 50  -> 64
    (2, 6, 6, 2, 1, 4),      # N = 72  -> 80
    (2, 6, 8, 2, 1, 3),      # N = 96
    (2, 7, 7, 3, 1, 4),      # N = 98  -> 112
    (2, 8, 8, 6, 1, 2),      # N = 128: one full M-tile, no padding rows
    (1, 12, 12, 2, 1, 3),    # N = 144: two M-tiles
    (2, 8, 10, 3, 1, 2),     # N = 160
    (2, 9, 9, 3, 2, 4),      # N = 162 -> 176
    (4, 6, 7, 2, 1, 3),      # N = 168 -> 176
    (2, 9, 9, 24, 3, 40),    # many (window, head) pairs per CTA: exercises the pipeline wrap-around
]


def _shift_regions(wd, wh, ww, nW):
    """Region ids shaped like sdf_window_index's: one cut per axis, present in some of the windows only."""
    dd, hh, wc = torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij")
    idx = torch.arange(nW).view(-1, 1)
    a, b, c = (idx % 2 == 1), (idx % 3 == 2), (idx % 4 >= 2)
    return (9 * a * (dd.reshape(1, -1) >= max(wd // 2, 1)) + 3 * b * (hh.reshape(1, -1) > wh // 2)
            + c * (wc.reshape(1, -1) > ww // 2)).to(torch.uint8).contiguous()


@pytest.mark.parametrize("case", FAST_CASES)
@pytest.mark.parametrize("mask", ["none", "shift", "random"])
@pytest.mark.parametrize("scale", [0.125, 32 ** -0.5])
def test_qktv_forward_fast_kernel(case, mask, scale):
    """The non-debug call (what the model issues) against the reference formula; for N <= 176 this is the
    warp-specialised kernel, which shares no epilogue code with the debug path tested above."""
    from sdformerflow_b200 import ops
    wd, wh, ww, nH, B, nW = case
    q, k, v, table, region, M, N, P, C = _inputs(wd, wh, ww, nH, B, nW, mask == "random", seed=7)
    if mask == "shift":
        region = _shift_regions(wd, wh, ww, nW)
    _, _, x = _reference(q, k, v, table, region, M, N, nH, (wd, wh, ww), scale)
    out, _, _ = ops.qktv_debug(q.to(DEV), k.to(DEV), v.to(DEV), table.to(DEV),
                               None if region is None else region.to(DEV), M, nH, nW, (wd, wh, ww), scale, debug=False)
    torch.cuda.synchronize()
    err = (out.cpu() - x).abs().max() / x.abs().max()
    assert err.item() <= 2e-6, err.item()


def test_qktv_forward_fast_kernel_is_deterministic():
    """The four warp roles hand over through mbarriers only; a missed hand-over shows up as run-to-run noise."""
    from sdformerflow_b200 import ops
    wd, wh, ww, nH, B, nW = 2, 9, 9, 3, 4, 60
    q, k, v, table, _, M, N, P, C = _inputs(wd, wh, ww, nH, B, nW, False, seed=11)
    region = _shift_regions(wd, wh, ww, nW).to(DEV)
    args = (q.to(DEV), k.to(DEV), v.to(DEV), table.to(DEV), region, M, nH, nW, (wd, wh, ww), 32 ** -0.5)
    first = ops.qktv_debug(*args, debug=False)[0].clone()
    for _ in range(20):
        again = ops.qktv_debug(*args, debug=False)[0]
        assert torch.equal(first, again)


@pytest.mark.parametrize("case", CASES)
def test_qktv_backward(case):
    from sdformerflow_b200 import ops
    wd, wh, ww, nH, B, nW, masked = case
    q, k, v, table, region, M, N, P, C = _inputs(*case, seed=3)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    tb = table.clone().requires_grad_(True)
    _, _, x = _reference(qf, kf, vf, tb, region, M, N, nH, (wd, wh, ww), 0.125)
    go = torch.randn(x.shape, generator=torch.Generator().manual_seed(5))
    x.backward(go)
    gq, gk, gv, gtab = ops.qktv_bwd_debug(q.to(DEV), k.to(DEV), v.to(DEV), table.to(DEV),
                                          None if region is None else region.to(DEV), go.to(DEV), M, nH, nW,
                                          (wd, wh, ww), 0.125)
    for got, ref, name in ((gq, qf.grad, "dq"), (gk, kf.grad, "dk"), (gv, vf.grad, "dv")):
        rel = (got.cpu().view_as(ref) - ref).norm() / ref.norm()
        assert rel.item() <= 5e-3, (name, rel.item())          # dO enters the MMAs in bf16
    rel = (gtab.cpu() - tb.grad).norm() / tb.grad.norm()
    assert rel.item() <= 5e-3, ("dtable", rel.item())


@pytest.mark.parametrize("variant", ["bn", "sdsa"])
@pytest.mark.parametrize("train", [False, True])
def test_sew_stage_against_reference_fixture(golden, variant, train):
    """Spiking_Swin_BasicLayer with Q K^T V attention (shifted + unshifted block, pad > 0) + SpikingPatchMerging."""
    from test_oracle_golden import sew_stage_inputs
    from sdformerflow_b200.STSwinNet_SNN import Spiking_swin_transformer3D as prod
    from sdformerflow_b200.sj import functional
    g = golden("sew_stage.pt")[f"{variant}_{'train' if train else 'eval'}"]
    P, x, cfg, spec = sew_stage_inputs(variant)
    kw = {"num_steps": 4, "v_reset": None, "v_th": 0.3, "neuron_type": "lif", "surrogate_fun": "surrogate.ATan()",
          "tau": 2.0, "detach_reset": True, "spike_norm": "BN"}

    class Blk(prod.Spiking_SwinTransformerBlock3D):
        attn_module = prod.Spiking_BN_WindowAttention3D if variant == "bn" else prod.SDSA_WindowAttention3D

    class Lyr(prod.Spiking_Swin_BasicLayer):
        swin_block_type = Blk
    lyr = Lyr(dim=64, input_resolution=(7, 10), depth=2, num_heads=2, window_size=(2, 3, 4),
              pretrained_window_size=(0, 0, 0), mlp_ratio=4.0, version="swinv1", qk_scale=0.125, drop_path=[0.0, 0.0],
              norm_layer="BN", downsample=prod.SpikingPatchMerging, **kw)
    lyr.load_state_dict({k[2:]: v for k, v in P.items()})
    lyr.train(train).to(DEV)
    # teacher-forced, block by block: each product block gets the oracle's input for that block
    import math
    xs = x.permute(0, 2, 3, 4, 1).contiguous()
    B, D, H, W, C = xs.shape
    shift_full = tuple(i // 2 for i in cfg.window_size)
    ws, ss = port.get_window_size((D, H, W), cfg.window_size, shift_full)
    Dp, Hp, Wp = (int(math.ceil(n / w)) * w for n, w in zip((D, H, W), ws))
    mask = port.compute_mask(Dp, Hp, Wp, ws, ss)
    worst = {}
    with torch.no_grad():
        for k in range(2):
            shift = (0, 0, 0) if k % 2 == 0 else shift_full
            ref = port.swin_block(xs, P, f"L.swin_blocks.{k}", cfg, cfg.num_heads[0], shift, mask, spec, port.BNMode(train))
            functional.reset_net(lyr)
            got = lyr.swin_blocks[k](xs.to(DEV)).cpu()
            worst[f"block{k}"] = ((got - ref).abs() > 1e-4 * ref.abs().max()).float().mean().item()
            xs = ref
        ref = port.patch_merging(xs, P, "L.downsample", cfg, spec, port.BNMode(train))
        functional.reset_net(lyr)
        got = lyr.downsample(xs.to(DEV)).cpu()
        worst["merging"] = ((got - ref).abs() > 1e-4 * ref.abs().max().clamp_min(1e-9)).float().mean().item()
    assert torch.allclose(xs, g["x_pre"], atol=2e-5)          # the oracle stream itself is the reference fixture
    print(f"SEW {variant} train={train}: teacher-forced mismatch fractions {worst}")
    for name, bad in worst.items():
        assert bad <= 2e-2, (name, bad)


@pytest.mark.parametrize("case", [c for c in FAST_CASES if c[4] * c[5] * c[3] <= 64])
@pytest.mark.parametrize("mask", ["none", "shift"])
def test_qktv_backward_fast_kernels(case, mask):
    """dQ, dK, dV and d(bias table) of the small-window kernels over every key-padding template, against autograd of
    the reference formula (dO enters the tensor cores in bf16, hence the 5e-3 relative tolerance, as for v1)."""
    from sdformerflow_b200 import ops
    wd, wh, ww, nH, B, nW = case
    q, k, v, table, _, M, N, P, C = _inputs(wd, wh, ww, nH, B, nW, False, seed=13)
    region = _shift_regions(wd, wh, ww, nW) if mask == "shift" else None
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    tb = table.clone().requires_grad_(True)
    scale = 32 ** -0.5
    _, _, x = _reference(qf, kf, vf, tb, region, M, N, nH, (wd, wh, ww), scale)
    go = torch.randn(x.shape, generator=torch.Generator().manual_seed(5))
    x.backward(go)
    gq, gk, gv, gtab = ops.qktv_bwd_debug(q.to(DEV), k.to(DEV), v.to(DEV), table.to(DEV),
                                          None if region is None else region.to(DEV), go.to(DEV), M, nH, nW,
                                          (wd, wh, ww), scale)
    for got, ref, name in ((gq, qf.grad, "dq"), (gk, kf.grad, "dk"), (gv, vf.grad, "dv")):
        rel = (got.cpu().view_as(ref) - ref).norm() / ref.norm()
        assert rel.item() <= 5e-3, (name, rel.item())
    rel = (gtab.cpu() - tb.grad).norm() / tb.grad.norm()
    assert rel.item() <= 5e-3, ("dtable", rel.item())
