"""sdformerflow_b200.train: the CUDA-graph training step and its data-parallel exchange.

  * a replayed GraphedStep is the same computation as the eager step (same losses, same weights after 3 steps);
  * parameters that never receive a gradient (PSN: dead attn_sn, reference Spiking_swin_transformer3D.py:711) are left
    alone, as plain AdamW in the reference's trainer leaves them;
  * 2 ranks (NCCL, needs 2 GPUs): after one data-parallel step every rank holds the same weights, and the all-reduced flat
    gradient equals the mean of the two single-GPU gradients on the same per-rank batches."""
import copy
import os
import socket

import pytest
import torch

from oracle import synth
from helpers import build_product

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inputs(B, seed=0):
    x = synth.synth_voxels(B, 10, 96, 128, seed=synth.SEED_INPUT + seed)
    gt, mask = synth.synth_labels(B, 96, 128, seed=synth.SEED_INPUT + 1 + seed)
    return x, gt, mask


def _model(nt="lif", device=DEV):
    mc, sc = synth.small_config(nt)
    model = build_product(mc, sc, device, train=True)
    for lyr in model.sttmultires_unet.encoders.swin3d.layers:          # no DropPath randomness
        for b in lyr.swin_blocks:
            if hasattr(b.drop_path, "forced"):
                b.drop_path.forced = torch.ones(2, device=device)
    return model


@pytest.mark.parametrize("nt", ["lif", "psn"])
def test_graphed_step_matches_eager(nt):
    from sdformerflow_b200 import train
    inp = [t.to(DEV) for t in _inputs(2)]
    runs = {}
    for graph in (False, True):
        model = _model(nt)
        state0 = copy.deepcopy(model.state_dict())
        step = train.GraphedStep(model, train.flow_loss, inp, lr=1e-3, weight_decay=0.01, graph=graph, clip_grad=100.0)
        with torch.no_grad():                          # construction warms up (and moves the weights): restart from state0
            for k, v in model.state_dict().items():
                v.copy_(state0[k])
            for st in step.opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        losses = [step(*inp).item()]
        after1 = {k: v.detach().clone() for k, v in model.named_parameters()}
        losses.append(step(*inp).item())
        runs[graph] = (losses, after1, step)
    (le, pe, _), (lg, pg, sg) = runs[False], runs[True]
    # step 1: the same computation from the same state (same loss, same updated weights).  Step 2 starts from weights that
    # differ wherever a ~0 gradient changed sign between the runs (atomics order in the bias / table reductions; Adam's first
    # step is +-lr whatever the magnitude) and a spiking net amplifies that: only its loss is compared, loosely.
    assert abs(le[0] - lg[0]) <= 1e-5 * abs(le[0]), (le, lg)
    assert abs(le[1] - lg[1]) <= 1e-2 * abs(le[1]), (le, lg)
    assert lg[1] != lg[0]                              # the weights really moved between replays
    # the captured step re-packs the weights up front on side streams (gemm.PackPlan): every Linear / conv parameter of the
    # spike GEMMs is in the plan, and nothing stays pinned after the capture
    from sdformerflow_b200 import gemm
    assert sg._plan is not None and len(sg._plan.jobs) >= 20 and not gemm._pinned
    n_bad = sum(((pe[k] - pg[k]).abs() > 2e-5).sum().item() for k in pe)
    assert n_bad <= 0.01 * sum(v.numel() for v in pe.values())
    n_params = sum(1 for p in sg.model.parameters() if p.requires_grad)
    if nt == "psn":
        assert len(sg.grads.params) < n_params        # dead attn_sn parameters are not in the optimizer
        dead = [n for n, p in sg.model.named_parameters() if p.grad is None]
        assert dead and all(".attn_sn." in n for n in dead), dead[:4]
    else:
        assert len(sg.grads.params) == n_params


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    from sdformerflow_b200 import train
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        model = _model("lif", dev)
        inp = [t.to(dev) for t in _inputs(2, seed=10 * rank)]          # a different batch per rank
        state0 = copy.deepcopy(model.state_dict())
        step = train.GraphedStep(model, train.flow_loss, inp, lr=1e-3, weight_decay=0.01, world=world, graph=True)
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(state0[k])
            for st in step.opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        step(*inp)
        torch.cuda.synchronize()
        # numpy arrays are pickled by value; a CPU tensor would travel as a shared-memory handle that dies with this process
        q.put((rank, step.grads.flat.detach().cpu().numpy(),
               torch.cat([p.detach().reshape(-1) for p in step.grads.params]).cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_rank_step_equals_mean_of_single_gpu_gradients():
    import torch.multiprocessing as mp
    from sdformerflow_b200 import train
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        r, flat, w = q.get(timeout=600)
        got[r] = (torch.from_numpy(flat), torch.from_numpy(w))
    for p in procs:
        p.join(timeout=60)
    assert torch.equal(got[0][0], got[1][0])           # both ranks hold the same reduced gradient ...
    assert torch.equal(got[0][1], got[1][1])           # ... and the same weights after the update
    # single-GPU gradients of the same two batches, averaged
    singles = []
    for r in range(2):
        model = _model("lif", DEV)
        inp = [t.to(DEV) for t in _inputs(2, seed=10 * r)]
        step = train.GraphedStep(model, train.flow_loss, inp, lr=1e-3, weight_decay=0.01, graph=False)
        model.load_state_dict(synth.synth_state_dict(model.state_dict(), 0))
        step._fwd_bwd()
        singles.append(step.grads.flat.detach().cpu().clone())
    mean = (singles[0] + singles[1]) / 2
    ref_scale = mean.abs().max()
    # identical kernels and inputs on both sides; only BN statistics order / atomics differ
    assert ((got[0][0] - mean).abs() > 1e-3 * ref_scale).float().mean().item() <= 1e-3
