"""GPU parity of the tcgen05 + TMA GEMM engine (csrc/spike_gemm.cu, csrc/spike_wgrad.cu) against fp64 references.

Forward (kind::i8 on u8 spikes x 3 weight digit planes): the result must be fp32-grade — the bar of the reference's fp32
Linear / Conv2d on spike tensors (Spiking_swin_transformer3D.py:126-131,267-290,909; Spiking_modules.py:268,318,803):
max |y - y64| <= 2e-6 * max|y64|.  It is also bit-reproducible: integer accumulation has no summation order.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gemm():
    from sdformerflow_b200 import gemm
    return gemm


def _spikes(shape, rate, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(shape, generator=g) < rate).to(torch.uint8).to(DEV)


# rows deliberately not multiples of 128; (K, Cout) = every Linear shape of the en4 model + odd ones
LINEAR_CASES = [
    (1000, 96, 192), (4096 + 77, 96, 96), (3000, 96, 384), (2500, 384, 96), (700, 384, 192), (1300, 192, 768),
    (900, 768, 192), (517, 768, 3072), (300, 3072, 768), (260, 1536, 768), (129, 768, 768), (5000, 96, 4), (640, 48, 48),
    (128, 16, 16),
]


@pytest.mark.parametrize("rows,K,Cout", LINEAR_CASES)
def test_spike_gemm_fwd_fp32_grade(rows, K, Cout):
    gemm = _gemm()
    torch.manual_seed(rows + K + Cout)
    a = _spikes((rows, K), 0.3, rows)
    w = (torch.randn(Cout, K) * 0.05).to(DEV)
    w[0, 0] = 0.37            # a large entry: the per-channel scale must cover the row maximum
    bias = torch.randn(Cout, device=DEV)
    pw = gemm.pack_weight(w)
    y, part = gemm.spike_gemm_fwd(a, pw, bias, want_stats=True)
    ref = a.double() @ w.double().t() + bias.double()
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 2e-6, err
    # BatchNorm partial sums emitted by the epilogue
    s = part[:, 0].double().sum(0)
    q = part[:, 1].double().sum(0)
    assert torch.allclose(s, y.double().sum(0), rtol=1e-5, atol=1e-3 * rows ** 0.5)
    assert torch.allclose(q, (y.double() ** 2).sum(0), rtol=1e-5, atol=1e-3)
    # bit-reproducible, with or without the statistics
    y2, _ = gemm.spike_gemm_fwd(a, pw, bias)
    assert torch.equal(y, y2)
    # spike operands (a_max = 1): conversion-free epilogue, at most 1 ulp (of the pre-bias value) from the exact path
    y3, part3 = gemm.spike_gemm_fwd(a, pw, bias, want_stats=True, a_max=1)
    err3 = (y3.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err3 <= 2e-6, err3
    assert (y3 - y).abs().max().item() <= 2.5e-7 * ref.abs().max().item()
    assert torch.equal(y3, gemm.spike_gemm_fwd(a, pw, bias, a_max=1)[0])
    assert torch.allclose(part3[:, 0].double().sum(0), y3.double().sum(0), rtol=1e-5, atol=1e-3 * rows ** 0.5)


def test_spike_gemm_integer_operand_and_exactness():
    """SEW residual sums are small integers: any u8 value is an exact operand; with weights that are exactly representable
    in 23-bit fixed point the GEMM is exact."""
    gemm = _gemm()
    g = torch.Generator().manual_seed(5)
    a = torch.randint(0, 200, (777, 192), generator=g, dtype=torch.uint8).to(DEV)
    w = (torch.randint(-2 ** 15, 2 ** 15, (96, 192), generator=g).float() / 2 ** 17).to(DEV)
    y, _ = gemm.spike_gemm_fwd(a, gemm.pack_weight(w))
    ref = a.double() @ w.double().t()
    assert torch.equal(y.double(), ref.float().double())


def test_pack_cache_never_serves_a_freed_parameters_planes():
    """A new parameter that recycles the id, the address and the version of a freed one (what the next model built by the
    same code does) must get its own digit planes, not the cached ones of the dead tensor."""
    gemm = _gemm()
    a = _spikes((256, 96), 0.3, 1)
    recycled = 0
    seen = set()
    for seed in range(8):
        torch.manual_seed(seed)
        w = torch.nn.Parameter(torch.randn(192, 96, device=DEV) * 0.05)
        recycled += (id(w), w.data_ptr()) in seen
        seen.add((id(w), w.data_ptr()))
        y, _ = gemm.spike_gemm_fwd(a, gemm.pack_weight(w), None, False, a_max=1)
        ref = a.double() @ w.detach().double().t()
        assert (y.double() - ref).abs().max().item() <= 2e-6 * ref.abs().max().item(), seed
        del w
    print("recycled (id, address) pairs:", recycled)


def test_pack_cache_follows_a_fused_optimizer_step():
    """torch's fused AdamW updates parameters WITHOUT bumping their _version: the pack caches are also keyed by an epoch that a
    global optimizer-step hook bumps, so the step after an update must use the new weights (Linear and transposed-conv packs)."""
    gemm = _gemm()
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(64, 32, device=DEV) * 0.1)
    wd = torch.nn.Parameter(torch.randn(32, 8, 3, 3, device=DEV) * 0.1)
    opt = torch.optim.AdamW([w, wd], lr=1e-1, fused=True)
    a = _spikes((200, 32), 0.3, 3)
    x = _spikes((1, 4, 4, 32), 0.3, 4)

    def check():
        y, _ = gemm.spike_gemm_fwd(a, gemm.pack_weight(w), None, False, a_max=1)
        ref = a.double() @ w.detach().double().t()
        assert (y.double() - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
        yd, _ = gemm.spike_deconv_fwd(x, gemm.pack_deconv_weight(wd), None, a_max=1)
        refd = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), wd.detach().double(), None, stride=2, padding=1,
                                  output_padding=1).permute(0, 2, 3, 1)
        assert (yd.double() - refd).abs().max().item() <= 2e-6 * refd.abs().max().item()

    check()
    assert gemm.pack_weight(w) is gemm.pack_weight(w)          # cached while nothing changes
    w.grad, wd.grad = torch.ones_like(w), torch.ones_like(wd)
    opt.step()
    check()


@pytest.mark.parametrize("rows,K,N", [(1000, 192, 96), (4099, 96, 96), (700, 768, 384), (513, 3072, 768), (300, 768, 3072),
                                     (2000, 384, 96), (200, 96, 48), (333, 4, 96), (260, 100, 20)])
def test_gemm_tf32(rows, K, N):
    gemm = _gemm()
    torch.manual_seed(rows + K)
    a = torch.randn(rows, K, device=DEV)
    b = torch.randn(N, K, device=DEV) * 0.05
    y = gemm.gemm_tf32(a, b)
    ref = a.double() @ b.double().t()
    # TF32 operands (10-bit mantissa, truncated) with fp32 accumulation
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 4e-3, err
    # fp32-grade when the operands are TF32-representable
    a2 = (torch.randint(-512, 512, (rows, K)).float() / 64).to(DEV)
    b2 = (torch.randint(-512, 512, (N, K)).float() / 1024).to(DEV)
    y2 = gemm.gemm_tf32(a2, b2)
    ref2 = a2.double() @ b2.double().t()
    # (fp32 accumulation of exact products: error ~ sqrt(K) * 2^-24)
    assert (y2.double() - ref2).abs().max().item() <= 2e-5 * ref2.abs().max().item()


CONV_CASES = [
    # Nimg, H, W, Cin, Cout, k, stride, pad
    (3, 24, 32, 96, 96, 3, 1, 1),
    (2, 20, 27, 96, 96, 3, 1, 1),       # partial patches in both directions
    (2, 48, 64, 48, 96, 3, 2, 1),       # patch-embed conv (stride 2)
    (2, 18, 24, 96, 96, 3, 2, 1),       # PED conv (stride 2, odd patch count)
    (5, 9, 12, 768, 768, 3, 1, 1),      # bottleneck res blocks
    (2, 16, 16, 96, 32, 1, 1, 0),       # 1x1
]


@pytest.mark.parametrize("Nimg,H,W,Cin,Cout,k,stride,pad", CONV_CASES)
def test_spike_conv_fwd_fp32_grade(Nimg, H, W, Cin, Cout, k, stride, pad):
    gemm = _gemm()
    torch.manual_seed(H * W + Cin)
    x = _spikes((Nimg, H, W, Cin), 0.25, H)
    w = (torch.randn(Cout, Cin, k, k) * 0.03).to(DEV)
    bias = torch.randn(Cout, device=DEV) * 0.1
    pw = gemm.pack_weight(w, "conv")
    y, part = gemm.spike_conv_fwd(x, pw, bias, k, k, stride, pad, want_stats=True, a_max=1)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 2e-6, err
    s = part[:, 0].double().sum(0)
    q = part[:, 1].double().sum(0)
    yy = y.double().reshape(-1, Cout)
    assert torch.allclose(s, yy.sum(0), rtol=1e-5, atol=1e-2)
    assert torch.allclose(q, (yy ** 2).sum(0), rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("rows,K,Cout", [(4099, 96, 384), (3000, 384, 96), (5000, 96, 192), (1300, 192, 768), (517, 768, 3072),
                                          (300, 3072, 768), (6480, 768, 768), (40, 96, 96), (2000, 1536, 768), (999, 48, 4)])
def test_spike_wgrad(rows, K, Cout):
    gemm = _gemm()
    torch.manual_seed(rows + K)
    s = _spikes((rows, K), 0.3, rows + 1)
    g = torch.randn(rows, Cout, device=DEV)
    dw = gemm.spike_wgrad(g, s)
    ref = g.double().t() @ s.double()
    err = (dw.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-4, err          # G enters as bf16 hi + bf16 lo (16 significant bits); spikes are exact
    assert torch.equal(dw, gemm.spike_wgrad(g, s, s_max=1))     # the 0/1 expansion produces the same bf16 operand
    s3 = (s * torch.randint(1, 4, s.shape, device=DEV, dtype=torch.uint8))       # general path: small integers
    ref3 = g.double().t() @ s3.double()
    assert (gemm.spike_wgrad(g, s3).double() - ref3).abs().max().item() <= 1e-4 * ref3.abs().max().item()
    # exact-operand case: fp32-grade
    g2 = (torch.randint(-512, 512, (rows, Cout)).float() / 256).to(DEV)
    dw2 = gemm.spike_wgrad(g2, s)
    ref2 = g2.double().t() @ s.double()
    assert (dw2.double() - ref2).abs().max().item() <= 2e-5 * ref2.abs().max().item()
    assert torch.equal(dw2, gemm.spike_wgrad(g2, s))            # fixed reduction order
    # bias gradient from the same pass over g
    dw3, db = gemm.spike_wgrad(g, s, want_db=True)
    assert torch.equal(dw3, dw)
    refb = g.double().sum(0)
    assert (db.double() - refb).abs().max().item() <= 1e-5 * g.double().abs().sum(0).max().item()
    assert torch.equal(db, gemm.spike_wgrad(g, s, want_db=True)[1])


@pytest.mark.parametrize("Nimg,H,W,Cin,Cout,k,stride,pad", CONV_CASES)
def test_spike_conv_wgrad(Nimg, H, W, Cin, Cout, k, stride, pad):
    gemm = _gemm()
    torch.manual_seed(H * W + Cin + 1)
    x = _spikes((Nimg, H, W, Cin), 0.25, H + 3)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    g = (torch.randint(-512, 512, (Nimg, Ho, Wo, Cout)).float() / 256).to(DEV)
    dw = gemm.spike_conv_wgrad(g, x, k, k, stride, pad)
    xr = x.permute(0, 3, 1, 2).double().requires_grad_(False)
    wr = torch.zeros(Cout, Cin, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xr, wr, None, stride=stride, padding=pad)
    (ref,) = torch.autograd.grad(y, wr, g.permute(0, 3, 1, 2).double())
    assert dw.shape == ref.shape
    assert (dw.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    dw3, db = gemm.spike_conv_wgrad(g, x, k, k, stride, pad, s_max=1, want_db=True)
    assert torch.equal(dw, dw3)
    refb = g.double().sum((0, 1, 2))
    assert (db.double() - refb).abs().max().item() <= 1e-5 * g.double().abs().sum((0, 1, 2)).max().item()


@pytest.mark.parametrize("Nimg,H,W,Cin,Cout", [(2, 9, 12, 1536, 384), (3, 18, 24, 784, 192), (2, 36, 48, 400, 96), (2, 20, 27, 208, 96),
                                                (1, 8, 16, 16, 4), (2, 5, 7, 32, 8)])
def test_spike_deconv_fwd_fp32_grade(Nimg, H, W, Cin, Cout):
    """ConvTranspose2d(3, stride 2, padding 1, output_padding 1) on spikes as four parity-class implicit GEMMs vs F.conv_transpose2d
    in fp64 (decoder shapes of the model incl. the channel counts padded to 16), BN partial sums of the four slabs."""
    gemm = _gemm()
    torch.manual_seed(H * W + Cin)
    x = _spikes((Nimg, H, W, Cin), 0.25, H + 5)
    w = (torch.randn(Cin, Cout, 3, 3) * 0.03).to(DEV)
    b = torch.randn(Cout).to(DEV)
    packs = gemm.pack_deconv_weight(w)
    y, part = gemm.spike_deconv_fwd(x, packs, b, want_stats=True, a_max=1)
    ref = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=2, padding=1,
                             output_padding=1).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert ((y.double() - ref).abs().max() / ref.abs().max()).item() <= 2e-6
    yy = y.double().reshape(-1, Cout)
    assert torch.allclose(part[:, 0].double().sum(0), yy.sum(0), rtol=1e-5, atol=1e-2)
    assert torch.allclose(part[:, 1].double().sum(0), (yy ** 2).sum(0), rtol=1e-5, atol=1e-2)
    # zero input slices appended by the packer (decoder inputs concatenated up to a multiple of 16 channels)
    if Cin >= 32:
        packs2 = gemm.pack_deconv_weight(w[:Cin - 14], cin=Cin)
        x2 = x.clone()
        x2[..., Cin - 14:] = 0
        y2, _ = gemm.spike_deconv_fwd(x2, packs2, b, a_max=1)
        ref2 = F.conv_transpose2d(x2[..., :Cin - 14].permute(0, 3, 1, 2).double(), w[:Cin - 14].double(), b.double(), stride=2,
                                  padding=1, output_padding=1).permute(0, 2, 3, 1)
        assert ((y2.double() - ref2).abs().max() / ref2.abs().max()).item() <= 2e-6


DECONV_BWD_CASES = [(2, 9, 12, 1536, 1536, 384), (2, 18, 24, 770, 784, 192), (1, 36, 48, 386, 400, 96), (2, 20, 27, 194, 208, 48),
                    (1, 8, 16, 16, 16, 4), (2, 5, 7, 18, 32, 8)]


@pytest.mark.parametrize("Nimg,H,W,Cin_w,Cin,Cout", DECONV_BWD_CASES)
def test_deconv_dgrad_tf32(Nimg, H, W, Cin_w, Cin, Cout):
    """Data gradient of ConvTranspose2d(3, stride 2, padding 1, output_padding 1) as ONE stride-2 TF32 implicit GEMM over g (decoder
    shapes incl. operands padded to 16 channels: the padding channels get zero gradients) vs fp64 autograd; operands are
    TF32-representable, so only the fp32 accumulation order differs."""
    gemm = _gemm()
    torch.manual_seed(H * W + Cin)
    g = (torch.randint(-512, 512, (Nimg, 2 * H, 2 * W, Cout)).float() / 256).to(DEV)
    w = (torch.randint(-512, 512, (Cin_w, Cout, 3, 3)).float() / 4096).to(DEV)
    dx = gemm.deconv_dgrad_tf32(g, w, Cin=Cin)
    xr = torch.zeros(Nimg, Cin_w, H, W, device=DEV, dtype=torch.float64, requires_grad=True)
    y = F.conv_transpose2d(xr, w.double(), None, stride=2, padding=1, output_padding=1)
    (ref,) = torch.autograd.grad(y, xr, g.permute(0, 3, 1, 2).double())
    ref = ref.permute(0, 2, 3, 1)
    assert dx.shape == (Nimg, H, W, Cin)
    assert (dx[..., :Cin_w].double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert not dx[..., Cin_w:].any()


@pytest.mark.parametrize("Nimg,H,W,Cin_w,Cin,Cout", DECONV_BWD_CASES)
def test_spike_deconv_wgrad(Nimg, H, W, Cin_w, Cin, Cout):
    """Weight + bias gradient of the transposed convolution as four parity-class G3 launches over strided views of g."""
    gemm = _gemm()
    torch.manual_seed(H * W + Cin + 2)
    x = _spikes((Nimg, H, W, Cin), 0.25, H + 7)
    x[..., Cin_w:] = 0
    g = (torch.randint(-512, 512, (Nimg, 2 * H, 2 * W, Cout)).float() / 256).to(DEV)
    dw, db = gemm.spike_deconv_wgrad(g, x, Cin_w=Cin_w, s_max=1, want_db=True)
    wr = torch.zeros(Cin_w, Cout, 3, 3, device=DEV, dtype=torch.float64, requires_grad=True)
    y = F.conv_transpose2d(x[..., :Cin_w].permute(0, 3, 1, 2).double(), wr, None, stride=2, padding=1, output_padding=1)
    (ref,) = torch.autograd.grad(y, wr, g.permute(0, 3, 1, 2).double())
    assert dw.shape == ref.shape
    assert (dw.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    refb = g.double().sum((0, 1, 2))
    assert (db.double() - refb).abs().max().item() <= 1e-5 * g.double().abs().sum((0, 1, 2)).max().item()
    dw2 = gemm.spike_deconv_wgrad(g, x, Cin_w=Cin_w, s_max=0)             # general byte expansion, no bias gradient
    assert torch.equal(dw, dw2)


@pytest.mark.parametrize("Nimg,H,W,Cin,Cout", [(3, 24, 32, 96, 96), (2, 21, 27, 48, 96), (2, 16, 17, 96, 32), (1, 144, 192, 96, 96)])
def test_conv_dgrad_s2_tf32(Nimg, H, W, Cin, Cout):
    """Data gradient of a 3x3 / stride-2 / padding-1 convolution as four parity-class TF32 implicit GEMMs (odd sizes included)."""
    gemm = _gemm()
    torch.manual_seed(H + W + Cin)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    g = (torch.randint(-512, 512, (Nimg, Ho, Wo, Cout)).float() / 256).to(DEV)
    w = (torch.randint(-512, 512, (Cout, Cin, 3, 3)).float() / 4096).to(DEV)
    dx = gemm.conv_dgrad_s2_tf32(g, w, H, W)
    xr = torch.zeros(Nimg, Cin, H, W, device=DEV, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xr, w.double(), None, stride=2, padding=1)
    (ref,) = torch.autograd.grad(y, xr, g.permute(0, 3, 1, 2).double())
    ref = ref.permute(0, 2, 3, 1)
    assert dx.shape == ref.shape
    assert (dx.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


@pytest.mark.parametrize("Nimg,H,W,Cin,Cout,k,pad", [(3, 24, 32, 96, 96, 3, 1), (2, 20, 27, 96, 96, 3, 1), (5, 9, 12, 768, 768, 3, 1),
                                                    (2, 16, 16, 48, 96, 1, 0)])
def test_conv_dgrad_tf32(Nimg, H, W, Cin, Cout, k, pad):
    gemm = _gemm()
    torch.manual_seed(H + W)
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    g = (torch.randint(-512, 512, (Nimg, Ho, Wo, Cout)).float() / 256).to(DEV)
    w = (torch.randint(-512, 512, (Cout, Cin, k, k)).float() / 4096).to(DEV)
    dx = gemm.conv_dgrad_tf32(g, w, H, W, pad)
    xr = torch.zeros(Nimg, Cin, H, W, device=DEV, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xr, w.double(), None, stride=1, padding=pad)
    (ref,) = torch.autograd.grad(y, xr, g.permute(0, 3, 1, 2).double())
    ref = ref.permute(0, 2, 3, 1)
    assert dx.shape == ref.shape
    assert (dx.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
