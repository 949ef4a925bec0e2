"""Kernel micro-bench sweep (BASELINE.json configs[4] / SURVEY.md §8d cfg5): LIF kernel T in {5,10,20}
on [T, N] fp32 with N = 8*120*160*384 (cfg2 stage-1 mlp.sn2), BN stats, QK-gate, window kernels.
Prints one JSON line per kernel: achieved GB/s (algorithmic bytes / CUDA-event time) vs the measured
HBM peak of MEASURED_PEAKS.json.  Inputs are far larger than L2 (126 MB)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdformerflow_b200 import ops, capi  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, nbytes, ms, **extra):
    pk, how = peak_gbs()
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algo_GB": round(nbytes / 1e9, 3), "GBps": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / pk, 3), "peak": pk, "peak_kind": how, **extra}), flush=True)


def main():
    dev = "cuda"
    N = 8 * 120 * 160 * 384
    cfg = ops.NeuronCfg(kind=capi.SDF_NEURON_LIF, v_th=0.1, v_reset=None, tau=2.0, detach_reset=True)
    for T in (5, 10, 20):
        x = torch.randn(T, N, device=dev) * 0.1 + 0.03
        lay = ops.seq_layout(x.shape, 0)
        for dt, nm, ob in ((capi.SDF_SPIKE_U8, "u8", 1), (capi.SDF_SPIKE_F32, "f32", 4)):
            ms = timeit(lambda: ops._lif_fwd_raw(x, lay, cfg.c(), dt))
            report(f"lif_fwd T={T} out={nm}", T * N * (4 + ob), ms, T=T, N=N)
        C = 384
        scale, shift = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
        ms = timeit(lambda: ops._lif_fwd_raw(x, lay, cfg.c(), capi.SDF_SPIKE_U8, scale, shift, C, 1))
        report(f"lif_fwd+bn T={T} out=u8", T * N * 5, ms, T=T, N=N)
        if T <= 20:          # T = 20: the vector path added in round 2 (u, h of 2 neurons x 20 steps in registers)
            gs = torch.randn(T, N, device=dev)
            gx = torch.empty_like(x)
            part = torch.empty(ops.N_PARTIAL, 2, C, device=dev)

            def bwd():
                capi.call("sdf_lif_bwd", capi.struct(
                    "sdf_lif_bwd_args", u=x.data_ptr(), grad_spike=gs.data_ptr(), grad_x=gx.data_ptr(),
                    scale=scale.data_ptr(), shift=shift.data_ptr(), bn_partials=part.data_ptr(),
                    n_partial_blocks=ops.N_PARTIAL, C=C, hw=1, lay=lay, neuron=cfg.c(),
                    stream=torch.cuda.current_stream().cuda_stream))
            ms = timeit(bwd)
            report(f"lif_bwd+bn T={T}", T * N * 12, ms, T=T, N=N)
            del gs, gx
        del x
        torch.cuda.empty_cache()
    # PSN (the shipped en4 configs' neuron, reference Spiking_submodules.py:196-211): forward, backward, parameter gradients
    T, Np = 10, 4 * 144 * 192 * 96
    xp = (torch.randn(T, Np, device=dev) * 0.2).requires_grad_(True)
    wp = (torch.eye(T, device=dev) + torch.randn(T, T, device=dev) * 0.05).requires_grad_(True)
    bp = torch.full((T, 1), -0.1, device=dev).requires_grad_(True)
    pcfg = ops.NeuronCfg(kind=capi.SDF_NEURON_LIF, v_th=0.1, v_reset=None, tau=2.0, detach_reset=True)   # surrogate settings only
    with torch.no_grad():
        ms = timeit(lambda: ops.psn(xp, wp, bp, pcfg, 0))
    report("psn_fwd T=10 out=f32", T * Np * 8, ms, T=T, N=Np)
    sp = ops.psn(xp, wp, bp, pcfg, 0)
    gsp = torch.randn_like(sp)
    ms = timeit(lambda: torch.autograd.grad(sp, (xp, wp, bp), gsp, retain_graph=True))
    report("psn_bwd T=10 (grad_x, dW / db accumulated in the same pass)", T * Np * 12, ms, T=T, N=Np)
    gh = torch.randn(T, Np, device=dev)
    ms = timeit(lambda: ops._psn_param_grads(gh, xp.detach()))
    report("psn_wgrad T=10 (dW [T,T], db; stand-alone kernel for layouts the fused path does not take)", T * Np * 8, ms, T=T, N=Np)
    del xp, sp, gsp, gh
    torch.cuda.empty_cache()
    # time-strided (B, D, H, W, C) layout, the MLP sn1 site of cfg2 stage 1
    x = torch.randn(8, 10, 120, 160, 96, device=dev) * 0.1
    lay = ops.seq_layout(x.shape, 1)
    ms = timeit(lambda: ops._lif_fwd_raw(x, lay, cfg.c(), capi.SDF_SPIKE_F32))
    report("lif_fwd (B,D,H,W,C) time-strided out=f32", x.numel() * 8, ms)
    # BN stats
    rows, C = 8 * 10 * 120 * 160, 384
    u = torch.randn(rows, C, device=dev)
    part = torch.empty(ops.N_PARTIAL, 2, C, device=dev)
    ms = timeit(lambda: capi.call("sdf_bn_stats", capi.struct(
        "sdf_bn_stats_args", x=u.data_ptr(), rows=rows, C=C, ld=C, partials=part.data_ptr(), n_partial_blocks=ops.N_PARTIAL,
        stream=torch.cuda.current_stream().cuda_stream)))
    report("bn_stats rows x 384", rows * C * 4, ms)
    del u
    # window LIF gather + QK-gate + scatter at cfg2 stage 1 (B=8, 120x160, C=96, window (2,9,9), shifted)
    B, D, H, W, C, nH = 8, 10, 120, 160, 96, 3
    geom = ops.WindowGeom.get(B, D, H, W, (2, 9, 9), (1, 4, 4), dev)
    x = torch.randn(B, D, H, W, C, device=dev) * 0.2
    ms = timeit(lambda: ops.lif_window_debug(x, geom, cfg)[0])
    report("lif_window_fwd (with h_seq)", x.numel() * 4 + geom.rows * C * 8, ms)
    qk = torch.randn(geom.rows, 2 * C, device=dev)
    sc, sh = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    pos = torch.randn(1, nH, geom.N, 32, device=dev) * 0.1
    gate = torch.empty(geom.rows, C, device=dev)

    def qkf():
        capi.call("sdf_attn_qkgate_fwd", capi.struct(
            "sdf_attn_qkgate_fwd_args", q_pre=qk.data_ptr(), k_pre=qk[:, C:].data_ptr(), ld=2 * C, q_scale=sc.data_ptr(),
            q_shift=sh.data_ptr(), k_scale=sc.data_ptr(), k_shift=sh.data_ptr(), pos=pos.data_ptr(), gate=gate.data_ptr(),
            wd=2, M=geom.M, P=geom.P, C=C, nH=nH, neuron=cfg.c(), spike_dtype=capi.SDF_SPIKE_F32,
            stream=torch.cuda.current_stream().cuda_stream))
    ms = timeit(qkf)
    report("attn_qkgate_fwd", geom.rows * C * 12, ms, rows=geom.rows)
    y = torch.randn(geom.rows, C, device=dev)
    ms = timeit(lambda: ops.window_scatter(y, geom, res=x))
    report("window_scatter+residual", geom.rows * C * 4 + x.numel() * 8, ms)
    # backward kernels of the same block
    gg = torch.randn(geom.rows, C, device=dev)
    gq, gk = torch.empty(geom.rows, C, device=dev), torch.empty(geom.rows, C, device=dev)
    pq, pk = torch.empty(ops.N_PARTIAL, 2, C, device=dev), torch.empty(ops.N_PARTIAL, 2, C, device=dev)

    def qkb():
        capi.call("sdf_attn_qkgate_bwd", capi.struct(
            "sdf_attn_qkgate_bwd_args", q_pre=qk.data_ptr(), k_pre=qk[:, C:].data_ptr(), ld=2 * C, q_scale=sc.data_ptr(),
            q_shift=sh.data_ptr(), k_scale=sc.data_ptr(), k_shift=sh.data_ptr(), pos=pos.data_ptr(), grad_gate=gg.data_ptr(),
            grad_q=gq.data_ptr(), grad_k=gk.data_ptr(), bn_partials_q=pq.data_ptr(), bn_partials_k=pk.data_ptr(),
            n_partial_blocks=ops.N_PARTIAL, wd=2, M=geom.M, P=geom.P, C=C, nH=nH, neuron=cfg.c(),
            stream=torch.cuda.current_stream().cuda_stream))
    ms = timeit(qkb)
    report("attn_qkgate_bwd (+BN partials)", geom.rows * C * 20, ms)
    coef = torch.randn(3, C, device=dev)
    du = torch.empty(geom.rows, C, device=dev)
    ms = timeit(lambda: capi.call("sdf_bn_bwd_apply", capi.struct(
        "sdf_bn_bwd_apply_args", dy=gg.data_ptr(), u=y.data_ptr(), ld_u=C, du=du.data_ptr(), ld_du=C, coef=coef.data_ptr(),
        rows=geom.rows, C=C, stream=torch.cuda.current_stream().cuda_stream)))
    report("bn_bwd_apply", geom.rows * C * 12, ms)
    gxw = torch.empty_like(x)
    ms = timeit(lambda: capi.call("sdf_lif_window_bwd", capi.struct(
        "sdf_lif_window_bwd_args", x=x.data_ptr(), grad_spike=gg.data_ptr(), grad_x=gxw.data_ptr(), win2x=geom.win2x.data_ptr(),
        wd=2, MP=geom.M * geom.P, C=C, neuron=cfg.c(), stream=torch.cuda.current_stream().cuda_stream)))
    report("lif_window_bwd", x.numel() * 8 + geom.rows * C * 4, ms)


def qktv():
    """K3/K4 sweep: algorithmic 4*N^2*32 FLOP per (window, pseudo-head) forward, 8*N^2*32 backward."""
    dev = "cuda"
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak = peaks.get("bf16_tflops", 1590.0)
    for (wd, wh, ww, C, nH, M, masked) in [(2, 9, 9, 96, 3, 10080, True), (2, 9, 9, 96, 3, 10080, False),
                                           (2, 9, 9, 768, 24, 240, True), (4, 12, 12, 96, 3, 3360, True)]:
        N, P = wd * wh * ww, wh * ww
        rows = wd * M * P
        g = torch.Generator(device=dev).manual_seed(0)
        q, k, v = ((torch.rand(rows, C, device=dev, generator=g) < 0.2).to(torch.uint8) for _ in range(3))
        table = torch.randn((2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), nH, device=dev) * 0.02
        nW = M // 8 if M % 8 == 0 else 1
        region = None
        if masked:
            # shifted-window region ids as sdf_window_index produces them: one cut per axis, only in the windows on the
            # far border of that axis (here: every 2nd window in depth, every 4th in height / width)
            dd, hh, wc = torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij")
            idx = torch.arange(nW).view(-1, 1)
            a, b, c = (idx % 2 == 1), ((idx // 2) % 4 == 3), ((idx // 8) % 4 == 3)
            region = (9 * a * (dd.reshape(1, -1) >= wd // 2) + 3 * b * (hh.reshape(1, -1) > wh // 2)
                      + c * (wc.reshape(1, -1) > ww // 2)).to(torch.uint8).to(dev).contiguous()
        out = torch.empty(rows, C, device=dev)

        def fwd():
            capi.call("sdf_attn_qktv_fwd", capi.struct(
                "sdf_attn_qktv_fwd_args", q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), bias_table=table.data_ptr(),
                region=None if region is None else region.data_ptr(), out=out.data_ptr(), M=M, nH=nH, nW=nW, wd=wd,
                wh=wh, ww=ww, scale=0.125, stream=torch.cuda.current_stream().cuda_stream))
        ms = timeit(fwd, iters=5)
        flop = 4.0 * N * N * 32 * M * nH
        nbytes = 3 * rows * C + 4 * rows * C
        pk, how = peak_gbs()
        print(json.dumps({"kernel": f"attn_qktv_fwd window=({wd},{wh},{ww}) C={C} M={M} mask={masked}", "ms": round(ms, 4),
                          "algo_TFLOPs": round(flop / ms / 1e9, 1), "frac_of_bf16_peak": round(flop / ms / 1e9 / tf_peak, 4),
                          "GBps": round(nbytes / ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ms / 1e6 / pk, 3)}), flush=True)
        go = torch.randn(rows, C, device=dev)
        gq, gk, gv = (torch.empty(rows, C, device=dev) for _ in range(3))
        gtab = torch.zeros_like(table)

        def bwd():
            capi.call("sdf_attn_qktv_bwd", capi.struct(
                "sdf_attn_qktv_bwd_args", q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), bias_table=table.data_ptr(),
                region=None if region is None else region.data_ptr(), grad_out=go.data_ptr(), grad_q=gq.data_ptr(),
                grad_k=gk.data_ptr(), grad_v=gv.data_ptr(), grad_bias_table=gtab.data_ptr(), M=M, nH=nH, nW=nW, wd=wd,
                wh=wh, ww=ww, scale=0.125, stream=torch.cuda.current_stream().cuda_stream))
        ms = timeit(bwd, iters=3)
        print(json.dumps({"kernel": f"attn_qktv_bwd window=({wd},{wh},{ww}) C={C} M={M}", "ms": round(ms, 4),
                          "algo_TFLOPs": round(2 * flop / ms / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    if "--qktv" in sys.argv:
        qktv()
    else:
        main()
        qktv()
