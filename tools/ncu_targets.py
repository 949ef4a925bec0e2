"""Launches each hot kernel a few times at its BASELINE.json size so that `ncu --set full -k regex:<name>` can
capture it in isolation (one GPU, no multi-rank).  Usage: python tools/ncu_targets.py [lif_fwd|lif_bwd|qkgate|qktv|all]"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdformerflow_b200 import ops, capi  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
cfg = ops.NeuronCfg(kind=capi.SDF_NEURON_LIF, v_th=0.1, v_reset=None, tau=2.0, detach_reset=True)
N, T, C = 8 * 120 * 160 * 384, 10, 384
if which in ("lif_fwd", "lif_bwd", "all"):
    x = torch.randn(T, N, device=dev) * 0.1 + 0.03
    lay = ops.seq_layout(x.shape, 0)
    scale, shift = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    if which in ("lif_fwd", "all"):
        for _ in range(3):
            ops._lif_fwd_raw(x, lay, cfg.c(), capi.SDF_SPIKE_U8, scale, shift, C, 1)
            ops._lif_fwd_raw(x, lay, cfg.c(), capi.SDF_SPIKE_F32, scale, shift, C, 1)
    if which in ("lif_bwd", "all"):
        gs, gx = torch.randn(T, N, device=dev), torch.empty(T, N, device=dev)
        part = torch.empty(ops.N_PARTIAL, 2, C, device=dev)
        for _ in range(3):
            capi.call("sdf_lif_bwd", capi.struct(
                "sdf_lif_bwd_args", u=x.data_ptr(), grad_spike=gs.data_ptr(), grad_x=gx.data_ptr(), scale=scale.data_ptr(),
                shift=shift.data_ptr(), bn_partials=part.data_ptr(), n_partial_blocks=ops.N_PARTIAL, C=C, hw=1, lay=lay,
                neuron=cfg.c(), stream=st()))
        del gs, gx
    del x
if which in ("qkgate", "all"):
    B, D, H, W, C2, nH = 8, 10, 120, 160, 96, 3
    geom = ops.WindowGeom.get(B, D, H, W, (2, 9, 9), (1, 4, 4), dev)
    qk = torch.randn(geom.rows, 2 * C2, device=dev)
    sc, sh = torch.ones(C2, device=dev), torch.zeros(C2, device=dev)
    pos = torch.randn(1, nH, geom.N, 32, device=dev) * 0.1
    gate = torch.empty(geom.rows, C2, device=dev)
    for _ in range(3):
        capi.call("sdf_attn_qkgate_fwd", capi.struct(
            "sdf_attn_qkgate_fwd_args", q_pre=qk.data_ptr(), k_pre=qk[:, C2:].data_ptr(), ld=2 * C2, q_scale=sc.data_ptr(),
            q_shift=sh.data_ptr(), k_scale=sc.data_ptr(), k_shift=sh.data_ptr(), pos=pos.data_ptr(), gate=gate.data_ptr(),
            wd=2, M=geom.M, P=geom.P, C=C2, nH=nH, neuron=cfg.c(), spike_dtype=capi.SDF_SPIKE_F32, stream=st()))
if which in ("qktv", "all"):
    for (wd, wh, ww, C3, nH, M) in [(2, 9, 9, 96, 3, 10080), (4, 12, 12, 96, 3, 3360)]:
        N3, P = wd * wh * ww, wh * ww
        rows = wd * M * P
        q, k, v = ((torch.rand(rows, C3, device=dev) < 0.2).to(torch.uint8) for _ in range(3))
        table = torch.randn((2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), nH, device=dev) * 0.02
        nW = M // 8
        dd, hh, wc = torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij")
        idx = torch.arange(nW).view(-1, 1)
        a_, b_, c_ = (idx % 2 == 1), ((idx // 2) % 4 == 3), ((idx // 8) % 4 == 3)
        region = (9 * a_ * (dd.reshape(1, -1) >= wd // 2) + 3 * b_ * (hh.reshape(1, -1) > wh // 2)
                  + c_ * (wc.reshape(1, -1) > ww // 2)).to(torch.uint8).to(dev).contiguous()
        out = torch.empty(rows, C3, device=dev)
        go = torch.randn(rows, C3, device=dev)
        gq, gk, gv = (torch.empty(rows, C3, device=dev) for _ in range(3))
        gtab = torch.zeros_like(table)
        # launch order (per window size): fwd masked x2, fwd unmasked x2, bwd masked x1 (= 3 kernels)
        for reg in (region, region, None, None):
            capi.call("sdf_attn_qktv_fwd", capi.struct(
                "sdf_attn_qktv_fwd_args", q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), bias_table=table.data_ptr(),
                region=None if reg is None else reg.data_ptr(), out=out.data_ptr(), M=M, nH=nH, nW=nW, wd=wd, wh=wh, ww=ww,
                scale=0.125, stream=st()))
        capi.call("sdf_attn_qktv_bwd", capi.struct(
            "sdf_attn_qktv_bwd_args", q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), bias_table=table.data_ptr(),
            region=region.data_ptr(), grad_out=go.data_ptr(), grad_q=gq.data_ptr(), grad_k=gk.data_ptr(), grad_v=gv.data_ptr(),
            grad_bias_table=gtab.data_ptr(), M=M, nH=nH, nW=nW, wd=wd, wh=wh, ww=ww, scale=0.125, stream=st()))
torch.cuda.synchronize()
print("done")
