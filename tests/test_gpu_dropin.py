"""GPU replay of what the reference's scripts do with the model, statement by statement, through the reference's own import
paths (sdformerflow_b200.dropin aliases) and the SHIPPED configuration values (tests/golden/ref_config_dsec_en4.json =
configs/train_DSEC_supervised_SDformerFlow_en4.yml parsed by the reference's configs/parser.py in the build container,
oracle/make_config_fixture.py).  The unmodified script itself is exercised on the CPU side (tests/test_dropin_cpu.py), where
the reference tree exists; here the same calls run for real on the B200:

  train_flow_parallel_supervised_SNN.py:62-73   input_size from the crop, eval(name)(model cfg, swin cfg), .to(device), init_weights
  :94-119                                        SG_alpha, reset_net, set_step_mode, neuron type lookup, set_backend("cupy", ...)
  :131-138                                       AdamW + MultiStepLR
  :236-239, :259-285, :299-336                   per batch: reset, pos/neg split, min-max norm, forward, loss, backward, clip, step
  eval_DSEC_flow_SNN.py:99,125,155,219           load_model (whole-module pickle -> state_dict -> strict=False), eval(), reset, forward
"""
import copy
import io
import json
import os

import pytest
import torch

from oracle import port, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _config():
    with open(os.path.join(GOLDEN, "ref_config_dsec_en4.json")) as f:
        return json.load(f)["config"]


def test_train_and_eval_script_sequence_shipped_config():
    from sdformerflow_b200 import dropin
    dropin.install()
    # the scripts' own import lines
    from models.STSwinNet_SNN.Spiking_STSwinNet import SpikingformerFlowNet, MS_SpikingformerFlowNet, MS_SpikingformerFlowNet_en4  # noqa: F401
    from models.STSwinNet_SNN.Spiking_submodules import PSN, GatedLIFNode, SLTTLIFNode  # noqa: F401  (import *)
    from spikingjelly.activation_based import functional, neuron, surrogate
    from torch.optim import AdamW  # noqa: F401  (from torch.optim import *)

    config = _config()
    device = torch.device(DEV)
    assert config["model"]["spiking_neuron"]["neuron_type"] == "psn" and config["swin_transformer"]["window_size"] == [2, 9, 9]
    config["swin_transformer"]["input_size"] = [config["loader"]["crop"][0], config["loader"]["crop"][1]]          # :62-63
    torch.manual_seed(config["loader"]["seed"])
    model = eval(config["model"]["name"])(config["model"].copy(), config["swin_transformer"].copy())                # :68
    model.to(device)                                                                                                  # :72
    model.init_weights()                                                                                              # :73
    if "SG_alpha" in config["optimizer"]:                                                                             # :94-97
        for m in model.modules():
            if isinstance(m, surrogate.ATan):
                m.alpha = config["optimizer"]["SG_alpha"]
    functional.reset_net(model)                                                                                       # :99
    functional.set_step_mode(model, config["data"]["step_mode"])                                                      # :100
    neurontype = {"if": getattr(neuron, "IFNode"), "lif": getattr(neuron, "LIFNode"), "plif": getattr(neuron, "ParametricLIFNode"),
                  "psn": PSN}[config["model"]["spiking_neuron"]["neuron_type"]]                                       # :103-116
    functional.set_backend(model, "cupy", neurontype)                                                                 # :118-119
    optimizer = eval(config["optimizer"]["name"])(model.parameters(), lr=config["optimizer"]["lr"],
                                                  weight_decay=config["optimizer"]["wd"])                             # :131-132
    scheduler = torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=config["optimizer"]["milestones"], gamma=0.5)
    optimizer.zero_grad()

    H, W = config["loader"]["crop"]
    B = 2
    g = torch.Generator().manual_seed(synth.SEED_INPUT)
    raw = torch.rand(B, config["model"]["num_bins"], H, W, generator=g) * (torch.rand(B, 10, H, W, generator=g) < 0.1) \
        * (torch.randint(0, 2, (B, 10, H, W), generator=g) * 2 - 1)
    label, mask = synth.synth_labels(B, H, W)
    losses = []
    model.train()
    for it in range(2):
        # second iteration: the `use_amp` branch of the script (:50, :248, :314-315, :329-331) — GradScaler + autocast around
        # forward and loss; the model switches autocast off for its own extent (fp32 membranes, exact integer spike GEMMs)
        scaler = torch.amp.GradScaler("cuda") if it == 1 else None
        functional.reset_net(model)                                                                                   # :238
        functional.set_step_mode(model, config["data"]["step_mode"])                                                  # :239
        chunk = raw.to(device=device, dtype=torch.float32)                                                            # :241
        lab, msk = label.to(device), mask.to(device)
        with torch.autocast("cuda", enabled=scaler is not None):                                                      # :248
            neg = torch.nn.functional.relu(-chunk)                                                                    # :261-265
            pos = torch.nn.functional.relu(chunk)
            chunk = torch.cat((torch.unsqueeze(pos, dim=2), torch.unsqueeze(neg, dim=2)), dim=2)
            mn, mx = torch.min(chunk[chunk != 0]), torch.max(chunk[chunk != 0])                                       # :278-284
            if not mn == mx:
                chunk[chunk != 0] = (chunk[chunk != 0] - mn) / (mx - mn)
            pred_list = model(chunk.to(device))                                                                       # :299
            pred = pred_list["flow"]                                                                                  # :300
            assert len(pred) == 4 and all(p.shape == (B, 2, H, W) and p.dtype == torch.float32 for p in pred)
            assert pred_list["attn"] is None
            loss = port.flow_loss(pred, lab, msk)                                                                     # :308 (loss/flow_supervised.py)
        assert torch.isfinite(loss)
        if scaler is not None:
            scaler.scale(loss).backward()                                                                             # :315
        else:
            loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), config["loss"]["clip_grad"])                               # :323-324
        if scaler is not None:
            scaler.step(optimizer)                                                                                    # :329-331
            scaler.update()
        else:
            optimizer.step()
        optimizer.zero_grad()
        losses.append(loss.item())
    scheduler.step()                                                                                                  # :488-489
    # with psn every neuron owns parameters; the dead attn_sn ones (reference :711) legitimately get no gradient
    assert losses[0] > 0 and losses[1] != losses[0]

    # ---- eval_DSEC_flow_SNN.py: checkpoint round trip (utils/utils.py:21-36,93-94) + inference ----
    buf = io.BytesIO()
    torch.save(model, buf)                                  # mlflow.pytorch.log_model(model, "model") writes this pickle
    buf.seek(0)
    pretrained_model = torch.load(buf, map_location=device, weights_only=False)                                       # :21
    pretrained_dict = {k.replace("module.", ""): v for k, v in pretrained_model.state_dict().items()}                 # :24-26
    model2 = eval(config["model"]["name"])(config["model"].copy(), config["swin_transformer"].copy()).to(device)      # eval :87-94
    res = model2.load_state_dict(pretrained_dict, strict=False)                                                       # :36
    assert not res.missing_keys and not res.unexpected_keys
    functional.reset_net(model2)
    functional.set_step_mode(model2, "m")
    functional.set_backend(model2, "cupy", neurontype)
    model.eval()
    model2.eval()                                                                                                     # eval :125
    with torch.no_grad():
        chunk = synth.synth_voxels(1, 10, H, W).to(device)
        functional.reset_net(model)
        functional.reset_net(model2)                                                                                  # eval :155
        f1 = model(chunk)["flow"][-1]                                                                                 # eval :219
        f2 = model2(chunk)["flow"][-1]
    assert torch.equal(f1, f2)
    assert torch.isfinite(f1).all()
