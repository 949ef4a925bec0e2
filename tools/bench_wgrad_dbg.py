"""wgrad kernel vs reduce kernel durations per layer shape (torch.profiler device times, no launch overhead).
SDF_WGRAD_DEBUG=1 skips the operand conversion, 2 also the MMAs, 3 runs the MMAs without TMA."""
import collections, json, os, sys, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
from sdformerflow_b200 import gemm
dev = "cuda"
lin = [("s1.fc1", 276480, 96, 384), ("s1.fc2", 276480, 384, 96), ("s2.fc1", 69120, 192, 768), ("s2.fc2", 69120, 768, 192),
       ("s3.qk", 17280, 384, 768), ("s3.proj", 17280, 384, 384), ("s3.fc1", 17280, 384, 1536), ("s3.fc2", 17280, 1536, 384),
       ("s4.qk", 4320, 768, 1536), ("s4.fc1", 4320, 768, 3072), ("s4.fc2", 4320, 3072, 768)]
conv = [("pe.res", 40, 144, 192, 96, 96, 1), ("pe.conv", 40, 288, 384, 48, 96, 2), ("pe.ped", 40, 144, 192, 96, 96, 2),
        ("bott.res", 40, 9, 12, 768, 768, 1)]
only = sys.argv[1:]
for name, *shape in lin + conv:
    if only and name not in only:
        continue
    if len(shape) == 3:
        rows, K, Cout = shape
        a8 = (torch.rand(rows, K, device=dev) < 0.3).to(torch.uint8); g = torch.randn(rows, Cout, device=dev)
        fn = lambda: gemm.spike_wgrad(g, a8, s_max=1)
    else:
        Nimg, H, W, Cin, Cout, stride = shape
        x8 = (torch.rand(Nimg, H, W, Cin, device=dev) < 0.2).to(torch.uint8)
        g = torch.randn(Nimg, H // stride, W // stride, Cout, device=dev)
        fn = lambda: gemm.spike_conv_wgrad(g, x8, 3, 3, stride, 1, s_max=1)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
    t = collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            t["reduce" if "reduce" in e.name else "wgrad" if "wgrad_kernel" in e.name else "other"] += e.device_time_total / 5
    print(json.dumps({"mode": os.environ.get("SDF_WGRAD_DEBUG", "0"), "case": name, "wgrad_us": round(t["wgrad"], 1),
                      "reduce_us": round(t["reduce"], 1), "other_us": round(t["other"], 1)}), flush=True)
