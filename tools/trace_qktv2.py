"""Debug aid: clock64 timeline of the K3 v2 roles (CTA 0).  Builds a -DSDF_V2_TRACE variant of the library
into _lib/libsdf_b200_trace.so, runs one (2,9,9) forward and prints per-item timestamps relative to the start."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdformerflow_b200 import build, capi

def build_trace_lib():
    out = os.path.join(build.OUT_DIR, "libsdf_b200_trace.so")
    srcs = [os.path.join(build.HERE, "csrc", s) for s in build.SOURCES]
    cmd = [build._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-DSDF_V2_TRACE", "-I", os.path.join(ROOT, "include"), "-shared", "-o", out] + srcs
    subprocess.check_call(cmd)
    return out

if __name__ == "__main__":
    if "--build" in sys.argv:
        print(build_trace_lib()); sys.exit(0)
    capi.LIB_PATH = os.path.join(build.OUT_DIR, "libsdf_b200_trace.so")
    import torch
    masked = "--mask" in sys.argv
    wd, wh, ww, C, nH, M = 2, 9, 9, 96, 3, 10080
    N, P = wd * wh * ww, wh * ww
    rows = wd * M * P
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    q, k, v = ((torch.rand(rows, C, device=dev, generator=g) < 0.2).to(torch.uint8) for _ in range(3))
    table = torch.randn((2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), nH, device=dev) * 0.02
    nW = M // 8
    region = None
    if masked:
        dd, hh, wc = torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij")
        idx = torch.arange(nW).view(-1, 1)
        a, b, c = (idx % 2 == 1), ((idx // 2) % 4 == 3), ((idx // 8) % 4 == 3)
        region = (9 * a * (dd.reshape(1, -1) >= wd // 2) + 3 * b * (hh.reshape(1, -1) > wh // 2)
                  + c * (wc.reshape(1, -1) > ww // 2)).to(torch.uint8).to(dev).contiguous()
    out = torch.empty(rows, C, device=dev)
    for _ in range(2):
        capi.call("sdf_attn_qktv_fwd", capi.struct(
            "sdf_attn_qktv_fwd_args", q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), bias_table=table.data_ptr(),
            region=None if region is None else region.data_ptr(), out=out.data_ptr(), M=M, nH=nH, nW=nW, wd=wd,
            wh=wh, ww=ww, scale=0.125, stream=torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    L = capi.lib()
    buf = (ctypes.c_longlong * (16 * 64))()
    L.sdf_debug_v2_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
    assert L.sdf_debug_v2_trace(buf) == 0
    t = [[buf[s * 64 + i] for i in range(64)] for s in range(16)]
    t0 = min(x for row in t for x in row if x > 0)
    names = {0: "M.loop", 3: "M.mma1issued", 1: "M.s16seen", 2: "M.mma2issued", 4: "E.wait_s", 5: "E.s_seen", 6: "E.s16_arrived", 7: "E.o_seen",
             8: "E.out_done", 10: "P.start", 11: "P.empty_seen", 12: "P.full_arrived", 13: "O.full_seen", 14: "O.rg_loaded", 15: "O.pass_done", 9: "O.combined"}
    print("item  " + "  ".join(f"{names[s]:>13}" for s in (0, 3, 1, 2, 4, 5, 6, 7, 8)))
    for i in range(30):
        print(f"{i:4d}  " + "  ".join(f"{(t[s][i] - t0 if t[s][i] else -1):13d}" for s in (0, 3, 1, 2, 4, 5, 6, 7, 8)))
    print("pair  " + "  ".join(f"{names[s]:>13}" for s in (10, 11, 12, 13, 14, 15, 9)))
    for i in range(8):
        print(f"{i:4d}  " + "  ".join(f"{(t[s][i] - t0 if t[s][i] else -1):13d}" for s in (10, 11, 12, 13, 14, 15, 9)))
