"""Drop-in boundary, script level: the reference's UNMODIFIED trainer runs through
``python -m sdformerflow_b200.dropin`` — its own imports (`models.STSwinNet_SNN.*`, `spikingjelly.activation_based.*`), its
real YAML through configs/parser.py, model construction by `eval(config["model"]["name"])`, init_weights, load_model,
reset_net / set_step_mode, AdamW + MultiStepLR, its DataLoader on a synthetic pre-processed dataset, RandomCrop / flips,
pos/neg split and min-max normalisation (train_flow_parallel_supervised_SNN.py:10-21,45-138,232-285) — up to its
`pred_list = model(chunk.to(device))` (:299), which lands in the B200 operators.  Without a GPU those raise (no CPU
fallback), which is what this test observes; tests/test_gpu_dropin.py replays the same sequence on the GPU.
Needs the reference tree (not present on the GPU box): skipped there."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not available")
def test_unmodified_train_script_reaches_hot_path(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU observation of the boundary; the GPU replay lives in test_gpu_dropin.py")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_harness", "run_ref_script.py"), str(tmp_path / "work"),
                        "train_flow_parallel_supervised_SNN.py", "--config", "configs/train_DSEC_supervised_SDformerFlow_en4.yml"],
                       capture_output=True, text=True, timeout=900)
    assert "DROPIN_REACHED_HOT_PATH" in r.stdout, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    # the script printed the model it built (train_flow_parallel_supervised_SNN.py:123): our classes, the shipped psn config
    assert "MS_SpikingformerFlowNet_en4" in r.stdout and "PSN" in r.stdout


def test_dropin_aliases_and_whole_module_pickle(tmp_path):
    """utils/utils.py:21-36 (load_model): torch.load of a whole-module pickle -> .state_dict() -> load_state_dict(strict=False).
    The pickle names classes by the reference's import paths; with the aliases installed they resolve to this package."""
    import copy
    import io
    import pickle
    import torch
    from oracle import synth
    from sdformerflow_b200 import dropin
    dropin.install()
    from models.STSwinNet_SNN.Spiking_STSwinNet import MS_SpikingformerFlowNet_en4, SpikingformerFlowNet  # noqa: F401
    from models.STSwinNet_SNN.Spiking_submodules import PSN, GatedLIFNode, SLTTLIFNode  # noqa: F401
    from spikingjelly.activation_based import functional, neuron, surrogate
    import sdformerflow_b200.STSwinNet_SNN.Spiking_STSwinNet as prod
    assert MS_SpikingformerFlowNet_en4 is prod.MS_SpikingformerFlowNet_en4
    mc, sc = synth.small_config("psn")
    import models.STSwinNet_SNN.Spiking_STSwinNet as ref_path_module
    Net = getattr(ref_path_module, mc["name"])                    # the scripts do eval(config["model"]["name"])
    model = Net(copy.deepcopy(mc), copy.deepcopy(sc))
    model.init_weights()
    functional.reset_net(model)
    functional.set_step_mode(model, "m")
    functional.set_backend(model, "cupy", PSN)
    functional.set_backend(model, "cupy", neuron.LIFNode)
    assert isinstance(surrogate.ATan(), torch.nn.Module)
    # Write the checkpoint the way the REFERENCE would have: every class named by the reference's import path
    # (what mlflow.pytorch.log_model pickles, utils/utils.py:93-94), none by this package's.
    import inspect
    import sdformerflow_b200.sj as sj_pkg
    renamed = []
    for alias, mod in list(sys.modules.items()):
        if alias.startswith(("models.STSwinNet_SNN.", "spikingjelly.activation_based.")):
            for _, cls in inspect.getmembers(mod, lambda o: inspect.isclass(o) or inspect.isfunction(o)):
                if cls.__module__ == mod.__name__ and cls.__module__.startswith("sdformerflow_b200"):
                    renamed.append((cls, cls.__module__))
                    cls.__module__ = alias
    try:
        buf = io.BytesIO()
        torch.save(model, buf, _use_new_zipfile_serialization=False)
    finally:
        for cls, orig in renamed:
            cls.__module__ = orig
    raw = buf.getvalue()
    import re
    # GLOBAL opcodes ("c<module>\n<name>\n", protocol 2): reference paths only
    assert re.search(rb"cmodels\.STSwinNet_SNN\.Spiking_STSwinNet\n", raw) and re.search(rb"cspikingjelly\.activation_based\.", raw)
    assert not re.search(rb"csdformerflow_b200[\w.]*\n", raw) and sj_pkg is not None
    loaded = torch.load(io.BytesIO(raw), map_location="cpu", weights_only=False)
    sd = {k.replace("module.", ""): v for k, v in loaded.state_dict().items()}
    fresh = Net(copy.deepcopy(mc), copy.deepcopy(sc))
    res = fresh.load_state_dict(sd, strict=False)
    assert not res.missing_keys and not res.unexpected_keys
    for (k1, v1), (k2, v2) in zip(model.state_dict().items(), fresh.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    # class lookup by the reference's path, as pickle does it (protocol 0 GLOBAL opcode)
    cls = pickle.loads(b"cmodels.STSwinNet_SNN.Spiking_STSwinNet\nMS_SpikingformerFlowNet_en4\n.")
    assert cls is prod.MS_SpikingformerFlowNet_en4
    assert pickle.loads(b"cspikingjelly.activation_based.neuron\nLIFNode\n.") is neuron.LIFNode
