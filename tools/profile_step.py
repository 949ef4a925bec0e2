"""Per-kernel time breakdown of one bench training step with torch.profiler (no replay, real clocks).
Guidance only — the judged launch list comes from ncu (profiles/)."""
import collections
import copy
import os
import re
import sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sdformerflow_b200.sj import functional  # noqa: E402
from sdformerflow_b200.STSwinNet_SNN import Spiking_STSwinNet as prod  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda")
mc, sc = bench.model_cfg()
for a in sys.argv[1:]:
    if a in ("lif", "psn", "plif"):
        mc["spiking_neuron"]["neuron_type"] = a
torch.manual_seed(0)
model = getattr(prod, mc["name"])(copy.deepcopy(mc), copy.deepcopy(sc))
model.init_weights()
model.to(dev).train()
functional.set_step_mode(model, "m")
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, fused=True)
x, gt, mask = (t.to(dev) for t in bench.synth_batch(4, 16146))


def step():
    functional.reset_net(model)
    loss = bench.flow_loss(model(x)["flow"], gt, mask)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
tot = collections.Counter()
cnt = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = re.sub(r"<.*", "", e.name)[:60]
        tot[k] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        cnt[k] += 1
T = sum(tot.values())
print(f"GPU busy {T / 2e3:.2f} ms/step")
for k, v in tot.most_common(45):
    print(f"{v / T * 100:6.2f}%  {v / 2e3:8.3f} ms/step  n={cnt[k] // 2:5d}  {k}")

if "--ops" in sys.argv:
    print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=28, max_name_column_width=40,
                                                             max_shapes_column_width=90))

if "--aten" in sys.argv:
    rows = [(e.self_device_time_total / 2e3, e.count // 2, e.key) for e in prof.key_averages() if e.key.startswith("aten::") or "Backward" in e.key or e.key.startswith("_")]
    for t, n, k in sorted(rows, reverse=True)[:40]:
        print(f"{t:8.3f} ms/step  n={n:5d}  {k}")

if "--shapes" in sys.argv:
    rows = [(e.self_device_time_total / 2e3, e.count // 2, e.key, str(e.input_shapes)[:110]) for e in prof.key_averages(group_by_input_shape=True)
            if e.key in ("aten::add_", "aten::add", "aten::sum", "aten::cat", "aten::copy_", "aten::mm", "aten::addmm", "aten::addmm_",
                         "aten::fill_", "aten::mul", "aten::convolution_backward", "aten::cudnn_convolution_transpose", "aten::cudnn_convolution")]
    for t, n, k, sh in sorted(rows, reverse=True)[:45]:
        print(f"{t:8.3f} ms/step  n={n:4d}  {k:34s} {sh}")
