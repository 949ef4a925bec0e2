import sys, torch
sys.path.insert(0, ".")
from sdformerflow_b200 import ops
from tools.bench_gemm import timeit
x = (torch.rand(40, 288, 384, 2, device="cuda") * (torch.rand(40, 288, 384, 2, device="cuda") < 0.1))
w = torch.randn(48, 2, 3, 3, device="cuda").requires_grad_(True); b = torch.zeros(48, device="cuda").requires_grad_(True)
with torch.no_grad():
    print("fwd ms", timeit(lambda: ops.conv3x3_small_cin(x, w, b)))
y = ops.conv3x3_small_cin(x, w, b); g = torch.randn_like(y)
print("bwd ms", timeit(lambda: torch.autograd.grad(y, (w, b), g, retain_graph=True)))
