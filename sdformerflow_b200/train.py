"""Training step of the reference's trainer, as the B200 runs it.

The reference's loop body (train_flow_parallel_supervised_SNN.py:236-336) is
    functional.reset_net(model); pred = model(chunk)["flow"]; loss = loss_function(...); loss.backward();
    clip_grad_norm_; optimizer.step(); optimizer.zero_grad()
single process, eager, under fp16 autocast with set_detect_anomaly(True) on every step (:236,:248).  SURVEY.md §8(f).3 asks
for the caller-side hygiene around the hot path; this module is that caller:

  * ``FlatGrads``      every parameter's .grad is a VIEW into one flat fp32 buffer: zero_grad is one memset, the data-
                       parallel exchange is ONE NCCL all-reduce on that buffer (no torch.cat / copy-back of 220 MB), and the
                       fused AdamW reads the same views.  Parameters that never receive a gradient (PSN / PLIF models: the
                       dead attn_sn parameters, reference Spiking_swin_transformer3D.py:711) are left without .grad, exactly as
                       in the reference, so AdamW skips them.
  * ``GraphedStep``    static shapes => reset + forward + loss + backward is captured once as a CUDA graph and replayed
                       (~1600 kernel launches per step become one graph launch); with one GPU the AdamW update is part of the
                       same graph, with several GPUs the sequence per step is
                           graph(reset, fwd, loss/world, bwd) -> all_reduce(flat grads, SUM) -> graph(AdamW)
                       i.e. the only eager launch is the collective.  BatchNorm statistics stay per replica, like the
                       reference's nn.DataParallel (train_mdr_supervised_SNN.py:125-128).
  * precision policy   the kernels integrate membranes in fp32 and run the spike GEMMs exactly (integer tensor-core
                       contraction) — there is nothing for autocast to speed up on this path, so the step runs in fp32 and
                       refuses to be entered under torch.autocast instead of silently ignoring it (the reference's
                       use_amp: True config key is accepted by ``fit`` and means "no GradScaler needed").
  * ``fit``            a small torchrun-ready epoch loop with the reference's optimizer / scheduler / clipping settings, for
                       users who do not drive the model from the reference's own script (sdformerflow_b200.dropin does that).
"""
import os

import torch
import torch.distributed as dist

from . import gemm
from .sj import functional


class FlatGrads:
    """Gradients of `params` as views of one contiguous buffer (see module docstring)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        # 16-byte aligned slices so that multi-tensor optimizer kernels keep their vector path
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, self.params)]
        self.attach()

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()

    def allreduce(self, group=None):
        """SUM over ranks (the loss is pre-divided by the world size, so this is the data-parallel mean)."""
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


def used_parameters(model, run_fwd_bwd):
    """Parameters that receive a gradient from one eager forward+backward (`run_fwd_bwd()` must call backward)."""
    for p in model.parameters():
        p.grad = None
    run_fwd_bwd()
    used = [p for p in model.parameters() if p.requires_grad and p.grad is not None]
    for p in model.parameters():
        p.grad = None
    return used


class GraphedStep:
    """One training step on static-shape inputs, replayed from CUDA graphs.

        step = GraphedStep(model, loss_fn, example=(x, gt, mask), lr=1e-4, weight_decay=0.01)
        loss = step(x, gt, mask)          # tensors on the model's device; returns the (static) loss tensor

    loss_fn(flows, gt, mask) -> scalar.  `world` > 1 expects an initialised NCCL process group."""

    def __init__(self, model, loss_fn, example, lr=1e-4, weight_decay=0.01, world=1, group=None, warmup=3, clip_grad=None,
                 graph=True, optimizer_cls=torch.optim.AdamW, prepack=True):
        if torch.is_autocast_enabled():
            raise RuntimeError("sdformerflow_b200.train: run the step in fp32 (membranes integrate in fp32, spike GEMMs are "
                               "exact integer contractions); torch.autocast is not supported on this path")
        self.model, self.loss_fn, self.world, self.group, self.clip_grad = model, loss_fn, world, group, clip_grad
        self.static = [t.clone() for t in example]
        self.graph_mode = bool(graph)

        def fwd_bwd_plain():
            functional.reset_net(model)
            loss = loss_fn(model(self.static[0])["flow"], *self.static[1:])
            loss.backward()

        self.loss = None
        self._fwd_graph = self._opt_graph = None
        self._plan = None
        # Everything that creates autograd state for the parameters (the probing pass, the flat gradient buffer, warm-up) runs
        # on ONE side stream, and the capture uses that same stream: autograd pins each AccumulateGrad node / .grad buffer to
        # the stream it was created on, and a captured backward must not have to synchronise with any other stream.
        self.stream = torch.cuda.Stream() if self.graph_mode else torch.cuda.current_stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            used = used_parameters(model, fwd_bwd_plain)
            self.grads = FlatGrads(used)
            self.opt = optimizer_cls(used, lr=lr, weight_decay=weight_decay, fused=True, capturable=self.graph_mode)
            if self.graph_mode:
                for _ in range(max(warmup, 3)):          # cudnn autotune, lazily built tables, optimizer state
                    self._eager()
                # which parameters the layers pack (digit planes for the spike GEMMs): recorded over one more eager step, then
                # re-packed at the START of every captured step on side streams instead of layer by layer (gemm.PackPlan)
                gemm.start_recording()
                try:
                    self._eager()
                finally:
                    jobs = gemm.stop_recording()
                self._plan = gemm.PackPlan(jobs) if (jobs and prepack) else None
        torch.cuda.current_stream().wait_stream(self.stream)
        torch.cuda.synchronize()
        if not self.graph_mode:
            return
        self._fwd_graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self._fwd_graph, stream=self.stream, capture_error_mode="thread_local"):
                self.loss = self._fwd_bwd()
                if world == 1:
                    self._update()
        finally:
            if self._plan is not None:
                self._plan.release()         # eager use of the model afterwards packs for itself
        if world > 1:
            self._opt_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._opt_graph, stream=self.stream, capture_error_mode="thread_local"):
                self._update()

    # -- pieces -------------------------------------------------------------------------------------------------------
    def _fwd_bwd(self):
        plan = self._plan if torch.cuda.is_current_stream_capturing() else None
        if plan is not None:
            plan.run()                       # forked branches of the graph; joined at the first layer that needs a pack
        self.grads.zero()
        functional.reset_net(self.model)
        loss = self.loss_fn(self.model(self.static[0])["flow"], *self.static[1:])
        if plan is not None:
            plan.join()
        (loss / self.world if self.world > 1 else loss).backward()
        return loss.detach()

    def _update(self):
        if self.clip_grad is not None:
            # reference :323-324 clip_grad_norm_(parameters, clip_grad); on the flat buffer it is one norm + one scale
            norm = torch.linalg.vector_norm(self.grads.flat)
            self.grads.flat.mul_(torch.clamp(self.clip_grad / (norm + 1e-6), max=1.0))
        self.opt.step()

    def _eager(self):
        loss = self._fwd_bwd()
        if self.world > 1:
            self.grads.allreduce(self.group)
        self._update()
        return loss

    # -- public -------------------------------------------------------------------------------------------------------
    def __call__(self, *inputs):
        for s, t in zip(self.static, inputs):
            if t is not s:
                s.copy_(t, non_blocking=True)
        if not self.graph_mode:
            self.loss = self._eager()
            return self.loss
        self._fwd_graph.replay()
        if self.world > 1:
            self.grads.allreduce(self.group)
            self._opt_graph.replay()
        gemm.bump_weights_epoch()       # the captured AdamW ran no Python hook: weight-derived caches are stale now
        return self.loss


def flow_loss(pred_list, gt, mask):
    """masked L2 end-point error averaged over the prediction scales — the reference's flow_loss_supervised with
    gamma None, lambda_mod 1 (loss/flow_supervised.py:14-31,81-105)."""
    nv = torch.sum(mask)
    cur = 0.0
    for pred in pred_list:
        err = torch.sqrt((pred - gt).pow(2).sum(1) + 1e-8).view(pred.shape[0], -1) * mask.reshape(pred.shape[0], -1)
        cur = cur + torch.sum(err, dim=1) / (nv + 1e-9)
    return torch.mean(cur / len(pred_list))


def init_distributed():
    """(rank, local_rank, world, device) from the torchrun environment; NCCL group when world > 1."""
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return rank, local_rank, world, dev


def fit(model, batches, config, epochs=1, loss_fn=flow_loss, log=print):
    """Epoch loop with the reference's settings (config = the dict its YAMLParser produces): AdamW(lr, wd) (:131-132),
    MultiStepLR(milestones, gamma 0.5) (:134-136), clip_grad (:323-324), one process per GPU, the batch of every rank taken
    from `batches(rank, world)` (an iterable of (chunk (B,bins,2,H,W), label, mask) on any device)."""
    rank, _local, world, dev = init_distributed()
    model.to(dev).train()
    functional.set_step_mode(model, config.get("data", {}).get("step_mode", "m"))
    oc = config["optimizer"]
    step = None
    history = []
    for epoch in range(epochs):
        total, n = 0.0, 0
        for chunk, label, mask in batches(rank, world):
            inp = tuple(t.to(dev, non_blocking=True).float() for t in (chunk, label, mask))
            if step is None:
                step = GraphedStep(model, loss_fn, inp, lr=oc["lr"], weight_decay=oc.get("wd", 0.01), world=world,
                                   clip_grad=config.get("loss", {}).get("clip_grad"))
                sched = torch.optim.lr_scheduler.MultiStepLR(step.opt, milestones=oc.get("milestones", []), gamma=0.5) \
                    if oc.get("scheduler") == "multistep" else None
            loss = step(*inp)
            total, n = total + float(loss), n + 1
        if step is not None and sched is not None:
            sched.step()
        history.append(total / max(n, 1))
        if rank == 0:
            log(f"epoch {epoch}: train_loss {history[-1]:.5f}")
    return history
