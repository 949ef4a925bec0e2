#!/usr/bin/env python
"""bench.py — headline benchmark of the SDformerFlow spiking Swin hot path on B200.

Metric (BASELINE.json): SDformerFlow forward+backward samples/s at 1/2/4/8 B200.
Workload (BASELINE.json configs[2]): MS_SpikingformerFlowNet_en4, lif neurons,
supervised training step = reset_net + forward + masked-EPE loss over 4 scales + backward
(surrogate gradient) + AdamW step, batch 4 per GPU, 288x384 crops of synthetic 10-bin DSEC-shaped
voxel grids, window (2,9,9); data parallel over N GPUs with a NCCL gradient all-reduce (DDP,
bucketed, overlapped with backward).  Weak scaling: per-GPU batch fixed.

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...             # the reference algorithm's CPU path (oracle port)

Prints ONE JSON line on rank 0.  `value` is timed with inputs resident in HBM; `e2e` repeats the
measurement through the public model API with pinned HOST buffers (H2D copy of the step's inputs
and a D2H read of the loss inside the timed region).  `roofline` reports the dominant libsdf_b200
entry point of the step (by summed CUDA-event time over an eager pass of the same step): algorithmic
bytes / event time of its launches against the measured HBM peak of MEASURED_PEAKS.json; `kernels`
lists every entry point the same way.  `cpu_baseline` is the oracle port timed on
the host cores in the same run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = "MS_SpikingformerFlowNet_en4 lif(v_th=0.1) train step fwd+loss+bwd+AdamW, B=4/GPU, 288x384, T=10, window (2,9,9)"
H, W, BINS, B_PER_GPU = 288, 384, 10, 4
METRIC = "SDformerFlow fwd+bwd samples/s"


def model_cfg():
    model = {
        "name": "MS_SpikingformerFlowNet_en4", "encoding": "voxel", "norm_input": "minmax", "num_bins": BINS,
        "base_num_channels": 96, "kernel_size": 3, "activations": ["relu", None], "final_activation": None,
        "mask_output": True, "norm": None, "use_upsample_conv": False,
        "spiking_neuron": {"num_steps": 10, "v_th": 0.1, "v_reset": None, "neuron_type": "lif",
                           "surrogate_fun": "surrogate.ATan()", "tau": 2.0, "detach_reset": True, "spike_norm": "BN"},
    }
    swin = {
        "use_arc": ["swinv1", "MS_PED_Spiking_PatchEmbed_Conv_sfn"], "state_combination": "none", "base_num_channels": 96,
        "swin_depths": [2, 2, 6, 2], "swin_num_heads": [3, 6, 12, 24], "swin_out_indices": [0, 1, 2, 3],
        "swin_patch_size": [1, 1, 2, 2], "window_size": [2, 9, 9], "pretrained_window_size": [0, 0, 0], "mlp_ratio": 4,
        "input_size": [H, W],
    }
    return model, swin


def synth_batch(B, seed):
    """voxels U(0,1)*[U(0,1)<0.10] (B,10,2,H,W); labels N(0,4^2); mask 1."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, BINS, 2, H, W, generator=g) * (torch.rand(B, BINS, 2, H, W, generator=g) < 0.10)
    gt = torch.randn(B, 2, H, W, generator=g) * 4.0
    mask = torch.ones(B, 1, H, W)
    return x, gt, mask


def flow_loss(pred_list, gt, mask):
    """masked L2-EPE averaged over the scales (reference loss/flow_supervised.py:14-31,81-105)."""
    nv = torch.sum(mask)
    cur = 0.0
    for pred in pred_list:
        err = torch.sqrt((pred - gt).pow(2).sum(1) + 1e-8).view(pred.shape[0], -1) * mask.reshape(pred.shape[0], -1)
        cur = cur + torch.sum(err, dim=1) / (nv + 1e-9)
    return torch.mean(cur / len(pred_list))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the reference's torch path, all host threads
# ---------------------------------------------------------------------------------------------
def cpu_train_step_factory(B):
    from oracle import port, synth
    from sdformerflow_b200.STSwinNet_SNN import Spiking_STSwinNet as prod
    import copy
    mc, sc = model_cfg()
    template = getattr(prod, mc["name"])(copy.deepcopy(mc), copy.deepcopy(sc)).state_dict()
    P = port.params_from_state_dict(synth.synth_state_dict(template, 0), requires_grad=True)
    params = [p for p in P.values() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01)
    cfg = port.FlowNetCfg(port.SwinCfg(window_size=sc["window_size"], depths=sc["swin_depths"],
                                       num_heads=sc["swin_num_heads"]), num_bins=BINS, num_steps=10)
    spec = port.NeuronSpec(10, "lif", 0.1, None, 2.0, True)
    x, gt, mask = synth_batch(B, 16146)
    scales = synth.synth_drop_scales(sc["swin_depths"], B)

    def step():
        flows = port.ms_flownet_forward(x, P, cfg, spec, port.BNMode(True), scales)
        loss = port.flow_loss(flows, gt, mask)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return float(loss)
    return step


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path.  /root/reference does not
    exist on the GPU box and the reference is pure Python over spikingjelly (not installable offline), so
    this arm times the oracle port (bit-exact w.r.t. the reference's modules, tests/test_oracle_golden.py)
    with every host thread.  One step = one B=1 training step of the same workload."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_train_step_factory(1)
    budget_s = 240.0
    t0 = time.time()
    step()
    t_first = time.time() - t0
    warm_done = 1
    while warm_done < args.warmup and (time.time() - t0) + t_first * (1 + 1) < budget_s * 0.4:
        step()
        warm_done += 1
    k_eff = max(1, min(args.steps, int((budget_s - (time.time() - t0)) / max(t_first, 1e-3))))
    t1 = time.time()
    for _ in range(k_eff):
        step()
    dt = time.time() - t1
    value = k_eff * 1 / dt
    sample = f"{k_eff} timed B=1 training steps (fwd+loss+bwd+AdamW) at 288x384 after {warm_done} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": k_eff,
        "steps_requested": args.steps, "warmup": warm_done, "ms_per_step": dt / k_eff * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_arm": "oracle port of the reference torch path on host CPU, B=1 per step"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_infer(args, rank, local_rank, world):
    """BASELINE.json configs[1]: MS_SpikingformerFlowNet_en4 inference, batch 8 per GPU, 10-bin 480x640, eval mode
    (BatchNorm folded into the neuron prologues), replicas only (no collective).  Extra line, not the headline."""
    import copy
    import torch.distributed as dist
    from sdformerflow_b200 import capi
    from sdformerflow_b200.sj import functional
    from sdformerflow_b200.STSwinNet_SNN import Spiking_STSwinNet as prod
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Hi, Wi, Bi = 480, 640, 8
    mc, sc = model_cfg()
    sc["input_size"] = [Hi, Wi]
    torch.manual_seed(0)
    model = getattr(prod, mc["name"])(copy.deepcopy(mc), copy.deepcopy(sc))
    model.init_weights()
    model.to(dev).eval()
    functional.set_step_mode(model, "m")
    g = torch.Generator().manual_seed(16146 + rank)
    xh = (torch.rand(Bi, BINS, 2, Hi, Wi, generator=g) * (torch.rand(Bi, BINS, 2, Hi, Wi, generator=g) < 0.10)).pin_memory()
    xd = xh.to(dev)

    def step(x):
        functional.reset_net(model)
        with torch.no_grad():
            return model(x)["flow"][-1]

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        step(xd)
    timer = capi.KernelTimer(only={"sdf_lif_fwd"})
    eager_step = step
    use_graph = args.graph in ("on", "auto")
    if use_graph:
        # one CUDA graph per forward (reset + model), replayed on a static input; K1 is timed in an eager pass below
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eager_step(xd)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        sx = xd.clone()
        graph = torch.cuda.CUDAGraph()
        n_cap = capi.launch_count()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            static_out = eager_step(sx)
        graph_launches = capi.launch_count() - n_cap

        def step(x):                     # noqa: F811
            if x is not sx:
                sx.copy_(x, non_blocking=True)
            graph.replay()
            return static_out
        xd = sx
        step(xd)
    else:
        capi.set_timer(timer)
    n0 = capi.launch_count()
    ms_total = timed(lambda: step(xd), args.steps)
    launches = capi.launch_count() - n0 if not use_graph else graph_launches * args.steps
    capi.set_timer(None)
    if use_graph:
        capi.set_timer(timer)
        timed(lambda: eager_step(xd), args.steps)
        capi.set_timer(None)
    ks = timer.summary().get("sdf_lif_fwd", {"launches": 0, "ms": 0.0, "bytes": 0, "gbps": 0.0})
    ms_e2e = timed(lambda: step(xh.to(dev, non_blocking=True)).sum().item(), args.steps)
    if rank == 0:
        peak, how = measured_peaks()
        print(json.dumps({
            "metric": "SDformerFlow inference samples/s", "value": world * Bi * args.steps / (ms_total * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "MS_SpikingformerFlowNet_en4 lif(v_th=0.1) inference (eval), B=8/GPU, 480x640, T=10, window (2,9,9)",
                       "parallelism": f"replicas x{world}",
                       "launch": "one CUDA graph per forward, replayed" if use_graph else "eager"},
            "e2e": {"value": world * Bi * args.steps / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": xh.numel() * 4,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "sdf_lif_fwd", "achieved": ks["gbps"], "peak": peak, "unit": "GB/s",
                         "frac": ks["gbps"] / peak, "traffic": None, "peak_kind": how, "launches": ks["launches"]},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _free_gpu():
    """Between workloads: collect Python cycles first (autograd nodes of eager warm-up steps can sit in reference cycles that
    pin their tensors), then hand the cached blocks back."""
    import gc
    gc.collect()
    torch.cuda.empty_cache()


def build_model(dev, Hh, Ww, neuron="lif", bins=BINS, window=(2, 9, 9), train=True, name=None):
    import copy
    from sdformerflow_b200.sj import functional
    from sdformerflow_b200.STSwinNet_SNN import Spiking_STSwinNet as prod
    mc, sc = model_cfg()
    mc["num_bins"] = bins
    mc["spiking_neuron"]["neuron_type"] = neuron
    mc["spiking_neuron"]["num_steps"] = bins
    sc["input_size"] = [Hh, Ww]
    sc["window_size"] = list(window)
    if name == "SpikingformerFlowNet":
        # the SEW-shortcut family (reference Spiking_STSwinNet.py:254-311): 3 encoder stages, Q K^T V window attention (K3/K4)
        mc["name"] = name
        sc["swin_depths"], sc["swin_num_heads"], sc["swin_out_indices"] = [2, 2, 6], [3, 6, 12], [0, 1, 2]
    torch.manual_seed(0)
    model = getattr(prod, mc["name"])(copy.deepcopy(mc), copy.deepcopy(sc))
    model.init_weights()
    model.to(dev).train(train)
    functional.set_step_mode(model, "m")
    functional.set_backend(model, "cupy", prod.neuron.LIFNode)   # accepted no-op, as the reference scripts call it
    return model


def synth_batch_shape(B, seed, bins, Hh, Ww):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, bins, 2, Hh, Ww, generator=g) * (torch.rand(B, bins, 2, Hh, Ww, generator=g) < 0.10)
    gt = torch.randn(B, 2, Hh, Ww, generator=g) * 4.0
    return x, gt, torch.ones(B, 1, Hh, Ww)


def make_timed(world, dev):
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()
    return timed


def train_workload(args, rank, local_rank, world, dev, Hh=H, Ww=W, B=B_PER_GPU, neuron="lif", bins=BINS, window=(2, 9, 9),
                   steps=None, full=True, name=None):
    """One training workload (reset + fwd + loss + bwd + AdamW, data parallel over `world` ranks) -> dict of measurements.
    full=True adds the eager pass with per-kernel events (roofline table) and the end-to-end (host buffers) measurement."""
    from sdformerflow_b200 import capi, train as sdtrain, distributed as sdist
    from sdformerflow_b200.sj import functional
    steps = steps or args.steps
    model = build_model(dev, Hh, Ww, neuron, bins, window, name=name)
    xh, gth, mh = synth_batch_shape(B, 16146 + rank, bins, Hh, Ww)
    xh, gth, mh = xh.pin_memory(), gth.pin_memory(), mh.pin_memory()
    xd, gtd, md = xh.to(dev), gth.to(dev), mh.to(dev)
    timed = make_timed(world, dev)
    use_graph = args.graph in ("on", "auto")
    out = {}
    if use_graph:
        # sdformerflow_b200.train.GraphedStep: flat gradient buffer, graph(reset+fwd+loss+bwd[+AdamW]); with N > 1 the only
        # eager launch per step is ONE NCCL all-reduce of the flat buffer between the two graphs
        stepper = sdtrain.GraphedStep(model, flow_loss, (xd, gtd, md), lr=1e-4, weight_decay=0.01, world=world,
                                      warmup=max(args.warmup, 3) + (5 if world > 1 else 0))
        n0 = capi.launch_count()
        stepper._eager()
        launches_per_step = capi.launch_count() - n0

        def step(x, gt, mask):
            return stepper(x, gt, mask)

        def eager_step(x, gt, mask):
            return stepper._eager()
        xd, gtd, md = stepper.static
    else:
        net = sdist.wrap(model, local_rank, find_unused_parameters=(neuron != "lif"))
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, fused=True)

        def step(x, gt, mask):
            functional.reset_net(model)
            loss = flow_loss(net(x)["flow"], gt, mask)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss
        eager_step = step
        launches_per_step = None
    for _ in range(max(args.warmup, 3)):
        step(xd, gtd, md)
    sampler = ClockSampler(local_rank)
    if rank == 0 and full:
        sampler.start()
    n0 = capi.launch_count()
    ms_total = timed(lambda: step(xd, gtd, md), steps)
    out["launches"] = (capi.launch_count() - n0) if launches_per_step is None else launches_per_step * steps
    out["clocks"] = sampler.stop() if (rank == 0 and full) else None
    out["ms_per_step"] = ms_total / steps
    out["value"] = world * B * steps / (ms_total * 1e-3)
    out["steps"] = steps
    if not full:
        return out
    # per-kernel CUDA events cannot be recorded inside a replay: an eager pass of the same step times every libsdf_b200
    # entry point (KernelTimer: CUDA events on the launching stream around each C-ABI call)
    timer = capi.KernelTimer()
    for _ in range(2):
        eager_step(xd, gtd, md)
    capi.set_timer(timer)
    ms_eager = timed(lambda: eager_step(xd, gtd, md), steps)
    capi.set_timer(None)
    out["eager_ms_per_step"] = ms_eager / steps
    out["kernels"] = {k: {"launches_per_step": v["launches"] / steps, "ms_per_step": v["ms"] / steps,
                          "algo_GB_per_step": v["bytes"] / steps / 1e9, "GBps": v["gbps"]}
                      for k, v in sorted(timer.summary().items(), key=lambda kv: -kv[1]["ms"])}

    def e2e_step():
        x = xh.to(dev, non_blocking=True)
        gt = gth.to(dev, non_blocking=True)
        mk = mh.to(dev, non_blocking=True)
        return step(x, gt, mk).item()
    e2e_step()
    ms_e2e = timed(e2e_step, steps)
    out["e2e_ms_per_step"] = ms_e2e / steps
    out["e2e_value"] = world * B * steps / (ms_e2e * 1e-3)
    out["h2d"] = (xh.numel() + gth.numel() + mh.numel()) * 4
    return out


def cpu_baseline_leg():
    """The oracle port (bit-exact restatement of the reference's torch path) on every host core: 1 warm-up + 3 timed B=1
    training steps of the headline workload (a B=4 step needs ~35 GB of host memory and ~4x the time; SURVEY.md §8d allows
    B=1 with the sample stated)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cstep = cpu_train_step_factory(1)
    cstep()
    t0 = time.time()
    n = 3
    for _ in range(n):
        cstep()
    dt = (time.time() - t0) / n
    return {"value": 1.0 / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{n} timed B=1 training steps (fwd+loss+bwd+AdamW) at 288x384 after 1 warm-up, oracle port of the "
                      "reference torch path (B=4 as on the GPU does not fit the time budget: 4x the work per step)"}


def committed_profile_numbers():
    """Numbers that only ncu can measure, read from the committed captures (profiles/): DRAM traffic of the dominant kernel and
    the tensor-pipe utilisation of the Q K^T V attention kernel (BASELINE.json metric 'attn TC util')."""
    out = {"traffic": None, "traffic_source": None, "attn_tc_util": None}
    p = os.path.join(ROOT, "profiles", "r02_ncu_numbers.json")
    if os.path.exists(p):
        out.update(json.load(open(p)))
    return out


def run_b200(args, rank, local_rank, world):
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.backends.cuda.matmul.allow_tf32 = False     # parity-grade fp32 for the few remaining library GEMMs / convs
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    head = train_workload(args, rank, local_rank, world, dev, neuron=args.neuron)
    secondary = {}
    if args.secondary == "on" or (args.secondary == "auto" and world == 1):
        # the other BASELINE.json configurations, short runs (3 timed steps each) in the same process
        k = 3
        _free_gpu()
        secondary["train_psn_288x384_shipped_neuron"] = train_workload(args, rank, local_rank, world, dev, neuron="psn", steps=k, full=False)
        _free_gpu()
        secondary["train_lif_480x640_B4"] = train_workload(args, rank, local_rank, world, dev, Hh=480, Ww=640, steps=k, full=False)
        _free_gpu()
        secondary["cfg4_train_T5_w288_256x256_B4"] = train_workload(args, rank, local_rank, world, dev, Hh=256, Ww=256, bins=5,
                                                                   window=(2, 8, 8), steps=k, full=False)
        _free_gpu()
        secondary["cfg4_train_T10_w466_192x192_B4"] = train_workload(args, rank, local_rank, world, dev, Hh=192, Ww=192,
                                                                    window=(4, 6, 6), steps=k, full=False)
        _free_gpu()
        secondary["infer_cfg2_B8_480x640"] = infer_workload(args, rank, world, dev, steps=k)
        secondary = {n: {"samples_per_s": v["value"], "ms_per_step": v["ms_per_step"], "steps": v["steps"]} for n, v in secondary.items()}
        _free_gpu()
        try:    # SEW family with the Q K^T V window attention (K3/K4) inside a whole training step; not a shipped config
            v = train_workload(args, rank, local_rank, world, dev, steps=k, full=False, name="SpikingformerFlowNet")
            secondary["train_sew_SpikingformerFlowNet_qktv_288x384_B4"] = {
                "samples_per_s": v["value"], "ms_per_step": v["ms_per_step"], "steps": v["steps"]}
        except Exception as e:  # noqa: BLE001 — a secondary line must not cost the headline
            secondary["train_sew_SpikingformerFlowNet_qktv_288x384_B4"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_leg()
    if rank == 0:
        peak, how = measured_peaks()
        prof = committed_profile_numbers()
        kernels = head["kernels"]
        top = next(iter(kernels))                      # dominant libsdf_b200 entry point by time in the step
        kt = kernels[top]
        lif = kernels.get("sdf_lif_fwd", {"GBps": 0.0, "ms_per_step": 0.0, "launches_per_step": 0})
        for v in kernels.values():
            v["frac_of_hbm_peak"] = v["GBps"] / peak
        own_ms = sum(v["ms_per_step"] for v in kernels.values())
        line = {
            "metric": METRIC, "value": head["value"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.replace("lif(v_th=0.1)", f"{args.neuron}(v_th=0.1)"), "global_batch": world * B_PER_GPU,
                       "parallelism": f"dp{world}",
                       "gemm": "own tcgen05 + TMA engine on every Linear / 3x3 conv fed by spikes: kind::i8 forward on 1-byte spikes x 3 "
                               "weight digit planes (exact integer accumulate, BN sums from the epilogue), TF32 data / weight "
                               "gradients; decoder transposed convs (forward as four parity-class implicit GEMMs, data gradient as a "
                               "stride-2 TF32 implicit GEMM, weight gradient as one G3 launch over four class tensor maps) and the strided "
                               "conv data gradient on the same engine; library (cuDNN/cuBLAS) only for real-valued operands (the 1x1 "
                               "stride-2 PED shortcut)",
                       "l2": "activations per step >> 126 MB L2; no explicit flush",
                       "weights": "random init (init_weights, seed 0)",
                       "launch": (("one CUDA graph per step (reset+fwd+loss+bwd+AdamW), replayed" if world == 1 else
                                   "graph(reset+fwd+loss+bwd) -> ONE NCCL all-reduce of the flat gradient buffer -> graph(AdamW)")
                                  + "; per-kernel roofline timed in an eager pass of the same step")
                       if args.graph != "off" else "eager (one launch per kernel), DDP for N > 1"},
            "e2e": {"value": head["e2e_value"], "unit": "samples/s", "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": 4,
                    "ms_per_step": head["e2e_ms_per_step"]},
            "gpu_launches": head["launches"],
            "clocks": head["clocks"],
            "eager": {"value": world * B_PER_GPU / (head["eager_ms_per_step"] * 1e-3), "unit": "samples/s",
                      "ms_per_step": head["eager_ms_per_step"],
                      "note": "same step launched kernel by kernel with per-kernel CUDA events on"},
            "roofline": {"bound": "hbm", "kernel": f"{top} (dominant libsdf_b200 entry point: {kt['ms_per_step']:.2f} ms of the step, "
                                                   f"{kt['launches_per_step']:.0f} launches; eager pass)",
                         "achieved": kt["GBps"], "peak": peak, "unit": "GB/s", "frac": kt["GBps"] / peak,
                         "traffic": (prof.get("traffic_by_entry") or {}).get(top, {}).get("dram_bytes_per_launch", prof["traffic"]),
                         "traffic_source": (prof.get("traffic_by_entry") or {}).get(top, {}).get("source", prof["traffic_source"]),
                         "peak_kind": how,
                         "algo_bytes_per_launch": kt["algo_GB_per_step"] * 1e9 / max(kt["launches_per_step"], 1),
                         "own_kernels_ms_per_step": own_ms,
                         "lif_fwd_K1": {"achieved": lif["GBps"], "frac": lif["GBps"] / peak, "ms_per_step": lif["ms_per_step"]}},
            "kernels": kernels,
            "attn_tc_util": prof["attn_tc_util"],
            "secondary": secondary,
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def infer_workload(args, rank, world, dev, steps):
    """BASELINE.json configs[1]: inference, B=8 per GPU, 480x640, eval (replicas only, no collective), one CUDA graph."""
    from sdformerflow_b200.sj import functional
    model = build_model(dev, 480, 640, "lif", train=False)
    g = torch.Generator().manual_seed(16146 + rank)
    xd = (torch.rand(8, BINS, 2, 480, 640, generator=g) * (torch.rand(8, BINS, 2, 480, 640, generator=g) < 0.10)).to(dev)

    def fwd():
        functional.reset_net(model)
        with torch.no_grad():
            return model(xd)["flow"][-1]
    for _ in range(3):
        fwd()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fwd()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        fwd()
    graph.replay()
    timed = make_timed(world, dev)
    ms = timed(graph.replay, steps)
    return {"value": world * 8 * steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the training step as a CUDA graph (auto = on); off = eager launches, DDP for N > 1")
    ap.add_argument("--neuron", default="lif", choices=["lif", "psn"],
                    help="neuron of the headline workload: lif (BASELINE.json north_star) or psn (the reference's shipped yml)")
    ap.add_argument("--secondary", default="auto", choices=["auto", "on", "off"],
                    help="also run the other BASELINE configs (psn, 480x640, cfg4, inference) as short secondary runs "
                         "(auto: only at N=1)")
    ap.add_argument("--workload", default="train", choices=["train", "infer"],
                    help="train (default, the headline metric: BASELINE.json configs[2]) or infer (configs[1]: eval, "
                         "B=8/GPU, 480x640)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    if args.workload == "infer":
        run_infer(args, rank, local_rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
