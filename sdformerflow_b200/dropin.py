"""Run the reference's scripts UNMODIFIED on the B200 implementation.

The reference scripts import the model and the spiking-neuron protocol by fixed paths
(train_flow_parallel_supervised_SNN.py:10,20-21; eval_DSEC_flow_SNN.py:6,14,16; train_mdr_supervised_SNN.py,
eval_MV_flow_SNN.py likewise):

    from models.STSwinNet_SNN.Spiking_STSwinNet import SpikingformerFlowNet, MS_SpikingformerFlowNet, MS_SpikingformerFlowNet_en4
    from models.STSwinNet_SNN.Spiking_submodules import *
    from spikingjelly.activation_based import functional, neuron, surrogate

and whole-module checkpoints written by mlflow (utils/utils.py:93-94, loaded at :21-36) pickle the classes under the same
paths.  ``install()`` registers this package's modules under those names in ``sys.modules`` — the rest of the reference
tree (configs/, loss/, utils/, the data loaders, the ``models`` package itself with its ANN models) is used as it is:

    cd SDformerFlow && python -m sdformerflow_b200.dropin train_flow_parallel_supervised_SNN.py --config configs/...yml
    cd SDformerFlow && python -m sdformerflow_b200.dropin eval_DSEC_flow_SNN.py --config configs/valid_DSEC_supervised.yml ...

or, inside a process:  ``import sdformerflow_b200.dropin as d; d.install()`` before the script's own imports.
"""
import importlib
import os
import runpy
import sys
import types

_MODEL_MODULES = ("Spiking_STSwinNet", "Spiking_submodules", "Spiking_modules", "Spiking_swin_transformer3D", "SNN_models")
_SJ_MODULES = ("functional", "neuron", "surrogate", "layer", "base")


def _package(name):
    """An importable (possibly pre-existing) package object for `name`; a stub package when nothing real exists."""
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        mod = types.ModuleType(name)
        mod.__path__ = []           # a package, so that `import name.sub` consults sys.modules
        sys.modules[name] = mod
        return mod


def install(spikingjelly=True):
    """Alias the reference's import paths to the sdformerflow_b200 modules.  Idempotent.

    spikingjelly=True also aliases ``spikingjelly.activation_based.{functional,neuron,surrogate,layer,base}`` to
    ``sdformerflow_b200.sj`` (replacing an installed spikingjelly for this process): the neuron classes the scripts hand to
    ``functional.set_backend`` and the classes named inside whole-module pickles are then the B200 ones."""
    from . import sj
    from .STSwinNet_SNN import (Spiking_STSwinNet, Spiking_submodules, Spiking_modules, Spiking_swin_transformer3D,
                                SNN_models)
    prod = dict(Spiking_STSwinNet=Spiking_STSwinNet, Spiking_submodules=Spiking_submodules, Spiking_modules=Spiking_modules,
                Spiking_swin_transformer3D=Spiking_swin_transformer3D, SNN_models=SNN_models)
    _package("models")
    pkg = _package("models.STSwinNet_SNN")
    for name in _MODEL_MODULES:
        sys.modules[f"models.STSwinNet_SNN.{name}"] = prod[name]
        setattr(pkg, name, prod[name])
    if spikingjelly:
        root = types.ModuleType("spikingjelly")
        root.__path__ = []
        ab = types.ModuleType("spikingjelly.activation_based")
        ab.__path__ = []
        root.activation_based = ab
        sys.modules["spikingjelly"] = root
        sys.modules["spikingjelly.activation_based"] = ab
        for name in _SJ_MODULES:
            mod = importlib.import_module(f"{sj.__name__}.{name}")
            sys.modules[f"spikingjelly.activation_based.{name}"] = mod
            setattr(ab, name, mod)
    return prod


def run_script(path, argv):
    """Execute a reference script as __main__ with the aliases installed (what `python script.py args` would do)."""
    path = os.path.abspath(path)
    sys.argv = [path] + list(argv)
    sys.path.insert(0, os.path.dirname(path))      # as `python script.py` does; the reference's own `models` package is found
    install()
    return runpy.run_path(path, run_name="__main__")


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m sdformerflow_b200.dropin <reference script.py> [script arguments]")
    run_script(sys.argv[1], sys.argv[2:])
