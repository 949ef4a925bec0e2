"""Times the tcgen05 GEMM engine against the library path it replaces (cuBLAS TF32 x 2) at the en4 / cfg3 shapes.
Run on the GPU box: python tools/bench_gemm.py > gpurun_out/r02_bench_gemm.jsonl"""
import json
import sys

import torch

sys.path.insert(0, ".")
from sdformerflow_b200 import gemm, ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = "cuda"
    # (name, rows, K, Cout) at cfg3: B=4, T=10, 288x384
    cases = [("s1.qk", 285120, 96, 192), ("s1.proj", 285120, 96, 96), ("s1.fc1", 276480, 96, 384), ("s1.fc2", 276480, 384, 96),
             ("s2.fc1", 69120, 192, 768), ("s2.fc2", 69120, 768, 192), ("s3.fc1", 17280, 384, 1536), ("s3.fc2", 17280, 1536, 384),
             ("s4.fc1", 4320, 768, 3072), ("s4.fc2", 4320, 3072, 768), ("merge1", 69120, 384, 192)]
    for name, rows, K, Cout in cases:
        a8 = (torch.rand(rows, K, device=dev) < 0.3).to(torch.uint8)
        af = a8.float()
        w = torch.randn(Cout, K, device=dev) * 0.05
        g = torch.randn(rows, Cout, device=dev)
        pw = gemm.pack_weight(w)
        wt = w.t().contiguous()
        t_fwd = timeit(lambda: gemm.spike_gemm_fwd(a8, pw, None, want_stats=True, a_max=1))
        t_fwd_ns = timeit(lambda: gemm.spike_gemm_fwd(a8, pw, None, want_stats=False, a_max=1))
        t_lib = timeit(lambda: ops._SpikeLinearFn.apply(af, w, None))
        t_dg = timeit(lambda: gemm.gemm_tf32(g, wt))
        torch.backends.cuda.matmul.allow_tf32 = True
        t_dg_lib = timeit(lambda: torch.mm(g, w))
        t_wg_lib = timeit(lambda: torch.mm(g.t(), af))
        torch.backends.cuda.matmul.allow_tf32 = False
        t_wg = timeit(lambda: gemm.spike_wgrad(g, a8))
        t_pack = timeit(lambda: gemm.pack_weight(w.clone()))
        fwd_bytes = rows * K + 4 * rows * Cout
        print(json.dumps({"case": name, "rows": rows, "K": K, "Cout": Cout,
                          "fwd_ms": round(t_fwd, 4), "fwd_nostats_ms": round(t_fwd_ns, 4), "fwd_lib_tf32x2_ms": round(t_lib, 4), "fwd_GBps": round(fwd_bytes / t_fwd / 1e6, 1),
                          "dgrad_ms": round(t_dg, 4), "dgrad_lib_ms": round(t_dg_lib, 4),
                          "dgrad_GBps": round(4 * rows * (K + Cout) / t_dg / 1e6, 1),
                          "wgrad_ms": round(t_wg, 4), "wgrad_lib_ms": round(t_wg_lib, 4),
                          "wgrad_GBps": round(rows * (4 * Cout + K) / t_wg / 1e6, 1), "pack_ms": round(t_pack, 4)}), flush=True)
    # convolutions: (name, Nimg, H, W, Cin, Cout, stride)
    import torch.nn.functional as F
    for name, Nimg, H, W, Cin, Cout, stride in [("pe.res", 40, 144, 192, 96, 96, 1), ("pe.conv", 40, 288, 384, 48, 96, 2),
                                                 ("pe.ped", 40, 144, 192, 96, 96, 2), ("bott.res", 40, 9, 12, 768, 768, 1)]:
        x8 = (torch.rand(Nimg, H, W, Cin, device=dev) < 0.2).to(torch.uint8)
        xf = x8.float().permute(0, 3, 1, 2)          # logical NCHW, channels_last strides
        w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.03
        pw = gemm.pack_weight(w, "conv")
        t_fwd = timeit(lambda: gemm.spike_conv_fwd(x8, pw, None, 3, 3, stride, 1, want_stats=True, a_max=1))
        t_lib = timeit(lambda: ops._SpikeConvFn.apply(xf, w, None, stride, 1, False, 0))
        y, _ = gemm.spike_conv_fwd(x8, pw, None, 3, 3, stride, 1)
        g = torch.randn_like(y)
        t_wg = timeit(lambda: gemm.spike_conv_wgrad(g, x8, 3, 3, stride, 1))
        t_dg = timeit(lambda: gemm.conv_dgrad_tf32(g, w, H, W, 1)) if stride == 1 else float("nan")
        flop = 2 * y.numel() * Cin * 9
        print(json.dumps({"case": name, "fwd_ms": round(t_fwd, 4), "fwd_lib_tf32x2_ms": round(t_lib, 4),
                          "fwd_TFLOPs": round(flop / t_fwd / 1e9, 1), "wgrad_ms": round(t_wg, 4),
                          "wgrad_TFLOPs": round(flop / t_wg / 1e9, 1),
                          "dgrad_ms": round(t_dg, 4)}), flush=True)


if __name__ == "__main__":
    main()
