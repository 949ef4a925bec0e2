"""One small launch of every hand-written tensor-core / TMA kernel family (and K1/K2/K5), for
    compute-sanitizer --tool memcheck  python tools/sanitizer_targets.py
    compute-sanitizer --tool racecheck python tools/sanitizer_targets.py
Logs go to profiles/r02_sanitizer_*.txt (tools/sanitize.sh)."""
import sys

import torch

sys.path.insert(0, ".")
from sdformerflow_b200 import gemm, ops, capi  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def spikes(*shape, rate=0.3):
    return (torch.rand(*shape, device=dev) < rate).to(torch.uint8)


# G1 / G2 / G3: linear
a = spikes(700, 96)
w = torch.randn(192, 96, device=dev) * 0.05
pw = gemm.pack_weight(w, need_wt=True)
y, part = gemm.spike_gemm_fwd(a, pw, None, want_stats=True, a_max=1)
g = torch.randn_like(y)
gemm.gemm_tf32(g, pw.wt)
gemm.spike_wgrad(g, a)
# conv (stride 1 and 2)
x = spikes(2, 20, 27, 96, rate=0.25)
wc = torch.randn(96, 96, 3, 3, device=dev) * 0.03
pc = gemm.pack_weight(wc, "conv", need_wt=True)
yc, _ = gemm.spike_conv_fwd(x, pc, None, 3, 3, 1, 1, want_stats=True, a_max=1)
gc = torch.randn_like(yc)
gemm.conv_dgrad_tf32(gc, wc, 20, 27, 1, pc.wt)
gemm.spike_conv_wgrad(gc, x, 3, 3, 1, 1)
y2, _ = gemm.spike_conv_fwd(x, pc, None, 3, 3, 2, 1)
gemm.spike_conv_wgrad(torch.randn_like(y2), x, 3, 3, 2, 1)
gemm.spike_wgrad(g, a, s_max=1, want_db=True)                       # 0/1 expansion + bias gradient
gemm.spike_conv_wgrad(gc, x, 3, 3, 1, 1, s_max=1, want_db=True)     # halo box (stride 1, one kernel row per tile)
# G1t: transposed convolution as four parity-class launches (208 = 194 channels padded to the TMA pitch)
xd = spikes(2, 9, 12, 208, rate=0.25)
wdc = torch.randn(208, 96, 3, 3, device=dev) * 0.03
gemm.spike_deconv_fwd(xd, gemm.pack_deconv_weight(wdc), None, want_stats=True, a_max=1)
# head convolution (2 real-valued channels): forward + dW / db
xv = torch.rand(3, 17, 23, 2, device=dev)
wh_ = (torch.randn(48, 2, 3, 3, device=dev) * 0.3).requires_grad_(True)
bh_ = torch.zeros(48, device=dev, requires_grad=True)
ops.conv3x3_small_cin(xv, wh_, bh_).sum().backward()
# PSN forward / backward with the parameter gradients accumulated in the same pass; T = 20 vector path of K2
pw_ = (torch.eye(10, device=dev) + 0.05 * torch.randn(10, 10, device=dev)).requires_grad_(True)
pb_ = torch.full((10, 1), -0.1, device=dev).requires_grad_(True)
up = torch.randn(10, 4096, device=dev, requires_grad=True)
ops.psn(up, pw_, pb_, ops.NeuronCfg(kind=capi.SDF_NEURON_LIF, v_th=0.1, v_reset=None, tau=2.0, detach_reset=True), 0).sum().backward()
u20 = torch.randn(20, 4096, device=dev, requires_grad=True)
ops.neuron(u20, ops.NeuronCfg(kind=capi.SDF_NEURON_LIF, v_th=0.1, v_reset=None, tau=2.0, detach_reset=True)).sum().backward()
# K3 / K4 (v2 pipeline, masked and unmasked) on a (2,3,4) window
wd, wh, ww, nH, M = 2, 3, 4, 3, 8
N = wd * wh * ww
q, k, v = (spikes(M * nH * N, 32, rate=0.2) for _ in range(3))
table = torch.randn((2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), nH, device=dev) * 0.02
region = torch.randint(0, 3, (4 * N,), device=dev, dtype=torch.uint8)
for reg, nW in ((None, 1), (region, 4)):
    out, _, _ = ops.qktv_debug(q, k, v, table, reg, M, nH, nW, (wd, wh, ww), 0.125, debug=False)
    ops.qktv_bwd_debug(q, k, v, table, reg, torch.randn_like(out), M, nH, nW, (wd, wh, ww), 0.125)
# K1 / K2
cfg = ops.NeuronCfg(kind=capi.SDF_NEURON_LIF, v_th=0.1, v_reset=None, tau=2.0, detach_reset=True)
u = torch.randn(10, 4096, device=dev, requires_grad=True)
s = ops.neuron(u, cfg)
s.sum().backward()
# backward of the transposed / strided convolutions (stride-2 TMA operand, parity-class tensor maps, strided output maps)
gd = torch.randn(2, 18, 24, 96, device=dev)
gemm.deconv_dgrad_tf32(gd, wdc[:194].contiguous(), Cin=208)
gemm.spike_deconv_wgrad(gd, xd, Cin_w=194, s_max=1, want_db=True)
gemm.conv_dgrad_s2_tf32(torch.randn_like(y2), wc, 20, 27, pc.wt)
# small feature map: 9 x 12 M tile (rows past the patch are never loaded / stored)
x9 = spikes(3, 9, 12, 96)
y9, _ = gemm.spike_conv_fwd(x9, pc, None, 3, 3, 1, 1, want_stats=True, a_max=1)
gemm.conv_dgrad_tf32(torch.randn_like(y9), wc, 9, 12, 1, pc.wt)
# BN-fused neurons in training mode: streamed K2 / PSN backward with the BN partial sums
bn = torch.nn.BatchNorm2d(96).to(dev).train()
ub = torch.randn(2, 10, 6, 8, 96, device=dev, requires_grad=True)
ops.bn_neuron(ub, bn, cfg, time_dim=1).sum().backward()
ub2 = torch.randn(2, 10, 6, 8, 96, device=dev, requires_grad=True)
import types  # noqa: E402
ops.bn_neuron(ub2, bn, cfg, time_dim=1, psn=types.SimpleNamespace(weight=pw_, bias=pb_)).sum().backward()
torch.cuda.synchronize()
print("sanitizer targets done")
