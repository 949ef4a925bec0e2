"""Test harness (not product): runs an UNMODIFIED reference script through sdformerflow_b200.dropin in a scratch working
directory that symlinks the reference tree and holds a tiny synthetic pre-processed DSEC dataset in the layout
DSEC_dataloader/DSEC_dataset_lite.py:36-136 reads.  Third-party packages the scripts import but this container lacks
(mlflow, matplotlib, imageio, h5py) get minimal stand-ins here — they are callers' logging / plotting / file libraries, not part
of the hot path.

    python tests/ref_harness/run_ref_script.py <workdir> train_flow_parallel_supervised_SNN.py [script args]

Prints DROPIN_REACHED_HOT_PATH when the script's own `model(chunk)` call reaches the B200 operators (on a machine without a
GPU those raise "no CPU fallback"; with a GPU the script simply trains).
"""
import os
import sys
import traceback
import types

import numpy as np

REF = os.environ.get("SDF_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_third_party_stand_ins(workdir):
    class _Info:
        artifact_uri = "file://" + os.path.join(workdir, "mlruns", "0", "run", "artifacts")
        run_id = "run"

    class _Run:
        info = _Info()
        data = types.SimpleNamespace(params={})

    def _get_run(run_id):
        raise RuntimeError("no such run")        # utils.load_model treats this as "no checkpoint" (utils/utils.py:11-14)

    noop = lambda *a, **k: None  # noqa: E731
    ml = _stub("mlflow", set_tracking_uri=noop, set_experiment=noop, start_run=lambda *a, **k: _Run(), end_run=noop,
               log_params=noop, log_param=noop, log_metric=noop, log_artifact=noop, active_run=lambda: _Run(), get_run=_get_run)
    ml.pytorch = _stub("mlflow.pytorch", log_model=noop, log_state_dict=noop)
    _stub("matplotlib", use=noop)
    sys.modules["matplotlib"].pyplot = _stub("matplotlib.pyplot")
    _stub("imageio")
    _stub("h5py")


def make_workdir(workdir, n_samples=2, H=480, W=640, bins=10):
    os.makedirs(workdir, exist_ok=True)
    for name in os.listdir(REF):
        dst = os.path.join(workdir, name)
        if not os.path.lexists(dst):
            os.symlink(os.path.join(REF, name), dst)
    base = os.path.join(workdir, "data", "Datasets", "DSEC", "saved_flow_data")
    rng = np.random.default_rng(16146)
    names = [f"synthetic_seq_a_{i:04d}.npy" for i in range(n_samples)]
    for sub in ("gt_tensors", "mask_tensors", os.path.join("event_tensors", f"{bins:02d}bins", "left", "synthetic_seq_a"),
                "sequence_lists"):
        os.makedirs(os.path.join(base, sub), exist_ok=True)
    for n in names:
        vox = rng.random((bins, H, W), dtype=np.float32) * (rng.random((bins, H, W)) < 0.1) * rng.choice([-1.0, 1.0], (bins, H, W))
        np.save(os.path.join(base, "event_tensors", f"{bins:02d}bins", "left", "synthetic_seq_a", n), vox.astype(np.float32))
        np.save(os.path.join(base, "gt_tensors", n), (rng.standard_normal((2, H, W)) * 4).astype(np.float32))
        np.save(os.path.join(base, "mask_tensors", n), np.ones((H, W), dtype=np.float32))
    for split in ("train", "valid"):
        with open(os.path.join(base, "sequence_lists", f"{split}_split_seq.csv"), "w") as f:
            f.write("\n".join(names) + "\n")


def main():
    workdir, script, argv = sys.argv[1], sys.argv[2], sys.argv[3:]
    sys.path.insert(0, ROOT)
    make_workdir(workdir)
    install_third_party_stand_ins(workdir)
    os.chdir(workdir)
    from sdformerflow_b200 import dropin
    try:
        dropin.run_script(os.path.join(workdir, script), argv)
    except RuntimeError as e:
        tb = traceback.extract_tb(e.__traceback__)
        frames = [(os.path.basename(f.filename), f.line) for f in tb]
        if "no CPU fallback" in str(e) and any(fn == script and "model(" in (ln or "") for fn, ln in frames):
            import torch
            from sdformerflow_b200.STSwinNet_SNN import Spiking_STSwinNet as prod
            print("DROPIN_REACHED_HOT_PATH", script, "model class module:", prod.__name__, "cuda:", torch.cuda.is_available())
            return
        raise


if __name__ == "__main__":
    main()
