"""Backward of the transposed / strided convolutions at the cfg3 shapes: the engine's kernels (sdf_deconv_dgrad_tf32,
sdf_spike_deconv_wgrad, sdf_conv_dgrad_s2_tf32) against the library path they replace (cuDNN TF32 on fp32-expanded spikes,
channels-last).  One JSON line per shape."""
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdformerflow_b200 import gemm  # noqa: E402
from tools.bench_gemm import timeit  # noqa: E402

torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
dev = "cuda"
NIMG = 40


def lib_conv_bwd(g4, x4, w, stride, transposed, mask):
    return torch.ops.aten.convolution_backward(g4, x4, w, None, [stride, stride], [1, 1], [1, 1], transposed,
                                               [1, 1] if transposed else [0, 0], 1, mask)


def main():
    # decoder transposed convs of the en4 model at 288x384, B*T = 40: (H, W, Cin_w, Cin padded to 16, Cout)
    for H, W, Cin_w, Cin, Cout in [(9, 12, 1536, 1536, 384), (18, 24, 770, 784, 192), (36, 48, 386, 400, 96), (72, 96, 194, 208, 48)]:
        x = (torch.rand(NIMG, H, W, Cin, device=dev) < 0.2).to(torch.uint8)
        x[..., Cin_w:] = 0
        g = torch.randn(NIMG, 2 * H, 2 * W, Cout, device=dev)
        w = torch.randn(Cin_w, Cout, 3, 3, device=dev) * 0.02
        wpad = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, Cin - Cin_w))
        g4, x4 = g.permute(0, 3, 1, 2), x.float().permute(0, 3, 1, 2)
        row = {"case": f"deconv {Cin_w}->{Cout} @{H}x{W}x{NIMG}",
               "own_dgrad_ms": timeit(lambda: gemm.deconv_dgrad_tf32(g, w, Cin=Cin)),
               "own_wgrad_db_ms": timeit(lambda: gemm.spike_deconv_wgrad(g, x, Cin_w=Cin_w, s_max=1, want_db=True)),
               "lib_dgrad_ms": timeit(lambda: lib_conv_bwd(g4, x4, wpad, 2, True, [True, False, False])),
               "lib_wgrad_ms": timeit(lambda: lib_conv_bwd(g4, x4, wpad, 2, True, [False, True, False])),
               "lib_expand_ms": timeit(lambda: x.float())}
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)
    # strided 3x3 convs of the patch embedding: (H, W, Cin, Cout)
    for H, W, Cin, Cout in [(288, 384, 48, 96), (144, 192, 96, 96)]:
        g = torch.randn(NIMG, H // 2, W // 2, Cout, device=dev)
        w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.02
        wd = w.permute(1, 2, 3, 0).reshape(Cin, 9 * Cout).contiguous()
        fake = g.new_empty((NIMG, H, W, Cin)).permute(0, 3, 1, 2)
        g4 = g.permute(0, 3, 1, 2)
        row = {"case": f"conv s2 {Cin}->{Cout} @{H}x{W}x{NIMG}",
               "own_dgrad_ms": timeit(lambda: gemm.conv_dgrad_s2_tf32(g, w, H, W, wd)),
               "lib_dgrad_ms": timeit(lambda: torch.ops.aten.convolution_backward(g4, fake, w, None, [2, 2], [1, 1], [1, 1], False, [0, 0], 1,
                                                                                 [True, False, False]))}
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)
    # regression guard for the stride-1 data gradient (res-block conv 96 -> 96 @144x192)
    g = torch.randn(NIMG, 144, 192, 96, device=dev)
    w = torch.randn(96, 96, 3, 3, device=dev) * 0.02
    wd = w.permute(1, 2, 3, 0).reshape(96, 9 * 96).contiguous()
    print(json.dumps({"case": "conv s1 dgrad 96->96 @144x192x40", "own_dgrad_ms": round(timeit(lambda: gemm.conv_dgrad_tf32(g, w, 144, 192, 1, wd)), 4)}))
    x = (torch.rand(NIMG, 144, 192, 96, device=dev) < 0.2).to(torch.uint8)
    pw = gemm.pack_weight(w, "conv", cache=False)
    print(json.dumps({"case": "conv s1 fwd 96->96 @144x192x40", "own_fwd_ms": round(timeit(lambda: gemm.spike_conv_fwd(x, pw, None, 3, 3, 1, 1, True, a_max=1)), 4)}))


if __name__ == "__main__":
    main()
