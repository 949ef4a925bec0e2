"""Per-stage parity of the product model (GPU) against the oracle port (CPU): where do they diverge?"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import port, synth  # noqa: E402
from helpers import port_spec, port_cfg, build_product  # noqa: E402
from sdformerflow_b200.sj import functional  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = "--tf32" in sys.argv

nt = "lif"
mc, sc = synth.small_config(nt)
model = build_product(mc, sc, "cuda", train=False)
x = synth.synth_voxels(2, 10, 96, 128)
P = port.params_from_state_dict(synth.synth_state_dict(model.state_dict(), 0))
P = {k: v.cpu() for k, v in P.items()}
feats, rec = [], None
with torch.no_grad():
    pe_ref = port.patch_embed_ms_ped(x, P, "sttmultires_unet.encoders.swin3d.patch_embed", port_spec(mc), port.BNMode(False), 10)
    flows_ref = port.ms_flownet_forward(x, P, port_cfg(mc, sc), port_spec(mc), port.BNMode(False), None, None, feats)
cap = {}
swin = model.sttmultires_unet.encoders.swin3d
swin.patch_embed.register_forward_hook(lambda m, i, o: cap.__setitem__("pe", o.detach().cpu()))
for i, lyr in enumerate(swin.layers):
    for k, blk in enumerate(lyr.swin_blocks):
        blk.register_forward_hook(lambda m, i_, o, key=(i, k): cap.__setitem__(key, o.detach().cpu()))
functional.reset_net(model)
with torch.no_grad():
    flows = model(x.cuda())["flow"]


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).abs() > 1e-4 * b.abs().max()).float().mean().item()


print("patch_embed", rel(cap["pe"], pe_ref))
# port per-block outputs: rerun stage by stage feeding the port's own stream
xs = pe_ref.permute(1, 0, 3, 4, 2).contiguous()
spec, mode, cfg = port_spec(mc), port.BNMode(False), port_cfg(mc, sc).swin
pre = "sttmultires_unet.encoders.swin3d"
for i in range(len(cfg.depths)):
    shift_full = tuple(s // 2 for s in cfg.window_size)
    for k in range(cfg.depths[i]):
        shift = (0, 0, 0) if k % 2 == 0 else shift_full
        with torch.no_grad():
            xs = port.swin_block(xs, P, f"{pre}.layers.{i}.swin_blocks.{k}", cfg, cfg.num_heads[i], shift, None, spec, mode)
        print("stage", i, "block", k, rel(cap[(i, k)], xs))
    if i < len(cfg.depths) - 1:
        with torch.no_grad():
            xs = port.patch_merging(xs, P, f"{pre}.layers.{i}.downsample", cfg, spec, mode)
for a, b in zip(flows, flows_ref):
    print("flow epe", (a.cpu() - b).pow(2).sum(1).sqrt().mean().item(), "ref mag", b.pow(2).sum(1).sqrt().mean().item())
