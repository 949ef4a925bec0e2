import json, sys, os, torch
sys.path.insert(0, ".")
from sdformerflow_b200 import gemm
from tools.bench_gemm import timeit
dev = "cuda"
x8 = (torch.rand(40, 144, 192, 96, device=dev) < 0.2).to(torch.uint8)
g = torch.randn(40, 144, 192, 96, device=dev)
t = timeit(lambda: gemm.spike_conv_wgrad(g, x8, 3, 3, 1, 1))
a8 = (torch.rand(276480, 96, device=dev) < 0.3).to(torch.uint8); gl = torch.randn(276480, 384, device=dev)
t2 = timeit(lambda: gemm.spike_wgrad(gl, a8))
a9 = (torch.rand(4320, 768, device=dev) < 0.3).to(torch.uint8); g9 = torch.randn(4320, 3072, device=dev)
t3 = timeit(lambda: gemm.spike_wgrad(g9, a9))
print(json.dumps({"mode": os.environ.get("SDF_WGRAD_DEBUG", "0"), "pe.res_ms": round(t, 4), "s1.fc1_ms": round(t2, 4), "s4.fc1_ms": round(t3, 4)}))
