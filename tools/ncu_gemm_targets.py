"""Single launches of the GEMM-engine kernels at bench sizes, for `ncu --set full -k regex:...` captures.
usage: python tools/ncu_gemm_targets.py {conv_wgrad|lin_wgrad|conv_fwd|lin_fwd|conv_dgrad|lin_dgrad|deconv_fwd}"""
import sys
import torch
sys.path.insert(0, ".")
from sdformerflow_b200 import gemm  # noqa: E402

dev = "cuda"
which = sys.argv[1]
torch.manual_seed(0)
if which == "deconv_fwd":
    x8 = (torch.rand(40, 72, 96, 208, device=dev) < 0.2).to(torch.uint8)       # last decoder: 194 channels padded to 208
    wt_ = torch.randn(208, 96, 3, 3, device=dev) * 0.03
    packs = gemm.pack_deconv_weight(wt_)
    fn = lambda: gemm.spike_deconv_fwd(x8, packs, None, want_stats=True, a_max=1)
elif which.startswith("conv"):
    x8 = (torch.rand(40, 144, 192, 96, device=dev) < 0.2).to(torch.uint8)
    w = torch.randn(96, 96, 3, 3, device=dev) * 0.03
    g = torch.randn(40, 144, 192, 96, device=dev)
    pw = gemm.pack_weight(w, "conv", need_wt=True)
    fn = {"conv_wgrad": lambda: gemm.spike_conv_wgrad(g, x8, 3, 3, 1, 1, s_max=1),
          "conv_fwd": lambda: gemm.spike_conv_fwd(x8, pw, None, 3, 3, 1, 1, want_stats=True, a_max=1),
          "conv_dgrad": lambda: gemm.conv_dgrad_tf32(g, w, 144, 192, 1, pw.wt)}[which]
else:
    a8 = (torch.rand(276480, 96, device=dev) < 0.3).to(torch.uint8)
    w = torch.randn(384, 96, device=dev) * 0.05
    g = torch.randn(276480, 384, device=dev)
    pw = gemm.pack_weight(w, need_wt=True)
    fn = {"lin_wgrad": lambda: gemm.spike_wgrad(g, a8, s_max=1),
          "lin_fwd": lambda: gemm.spike_gemm_fwd(a8, pw, None, want_stats=True, a_max=1),
          "lin_dgrad": lambda: gemm.gemm_tf32(g, pw.wt)}[which]
for _ in range(3):
    fn()
torch.cuda.synchronize()
