"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference model code.

Puts ``oracle/standin`` (spikingjelly/timm stand-ins) and
``/root/reference`` on ``sys.path`` and builds reference models the way the reference's
scripts do (train_flow_parallel_supervised_SNN.py:68-73, 99-100).  ``/root/reference``
only exists in the build container, so this module is used exclusively by
``oracle/make_golden.py`` and by CPU tests that skip when the tree is absent; nothing that
runs on the GPU box may import it.
"""
import copy
import os
import sys

REFERENCE_ROOT = os.environ.get("SDF_REFERENCE_ROOT", "/root/reference")
_STANDIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standin")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "STSwinNet_SNN"))


def _activate():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_STANDIN, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # sdformerflow_b200.dropin.install() registers the product under the reference's import path (``models``,
    # ``spikingjelly``) in sys.modules; if a test did that earlier in this process, drop those aliases so that the
    # reference's own package is what gets imported here.
    for top, home in (("models", REFERENCE_ROOT), ("spikingjelly", _STANDIN)):
        m = sys.modules.get(top)
        if m is not None and not (getattr(m, "__file__", None) or "").startswith(home):
            for k in [k for k in sys.modules if k == top or k.startswith(top + ".")]:
                del sys.modules[k]


def default_config(neuron_type="lif", v_th=0.1, num_steps=10, window_size=(2, 9, 9),
                   input_size=(288, 384), swin_depths=(2, 2, 6, 2), swin_num_heads=(3, 6, 12, 24),
                   base_num_channels=96, num_bins=10, name="MS_SpikingformerFlowNet_en4"):
    """Config dicts shaped like YAMLParser + combine_entries produce them
    (configs/train_DSEC_supervised_SDformerFlow_en4.yml, configs/parser.py:123-133)."""
    n = len(swin_depths)
    model = {
        "name": name, "encoding": "voxel", "norm_input": "minmax", "num_bins": num_bins,
        "base_num_channels": base_num_channels, "kernel_size": 3, "activations": ["relu", None],
        "final_activation": None, "mask_output": True, "norm": None, "use_upsample_conv": False,
        "spiking_neuron": {
            "num_steps": num_steps, "v_th": v_th, "v_reset": None, "neuron_type": neuron_type,
            "surrogate_fun": "surrogate.ATan()", "tau": 2., "detach_reset": True, "spike_norm": "BN",
        },
    }
    swin = {
        "use_arc": ["swinv1", "MS_PED_Spiking_PatchEmbed_Conv_sfn"], "state_combination": "none",
        "base_num_channels": base_num_channels, "swin_depths": list(swin_depths),
        "swin_num_heads": list(swin_num_heads), "swin_out_indices": list(range(n)),
        "swin_patch_size": [1, 1, 2, 2], "window_size": list(window_size),
        "pretrained_window_size": [0, 0, 0], "mlp_ratio": 4, "input_size": list(input_size),
    }
    return model, swin


def build_reference_model(model_cfg, swin_cfg, seed=0, train=False):
    """eval(config.model.name)(config.model.copy(), config.swin_transformer.copy()) + init_weights
    + reset_net + set_step_mode('m'), as the reference scripts do."""
    _activate()
    import torch
    from models.STSwinNet_SNN import Spiking_STSwinNet as ref
    from spikingjelly.activation_based import functional
    cls = getattr(ref, model_cfg["name"])
    torch.manual_seed(seed)
    model = cls(copy.deepcopy(model_cfg), copy.deepcopy(swin_cfg))
    model.init_weights()
    functional.reset_net(model)
    functional.set_step_mode(model, "m")
    model.train(train)
    return model


def reference_modules():
    """Returns the reference's hot-path python modules (swin, modules, submodules)."""
    _activate()
    from models.STSwinNet_SNN import Spiking_swin_transformer3D as swin
    from models.STSwinNet_SNN import Spiking_modules as mods
    from models.STSwinNet_SNN import Spiking_submodules as sub
    from spikingjelly.activation_based import functional
    return swin, mods, sub, functional
