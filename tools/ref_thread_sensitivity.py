"""Evidence for DESIGN.md §4 (the "chaos" claim), CPU only, build container only (needs /root/reference):
the UNMODIFIED reference model, same weights, same input, evaluated twice — with 1 CPU thread and with N threads.  The only
difference between the two runs is the summation order inside the library GEMM / conv / BatchNorm reductions (how MKL /
oneDNN split the work), i.e. fp32 rounding at the 1e-7 level.  If the reference's OWN end-to-end flow moves by far more than
1e-3 px between the two, then "flow within 1e-3 px free-running" is not a property any re-implementation with a different
summation order can have, and the attainable gate is "within the reference's own thread-count sensitivity".

    python tools/ref_thread_sensitivity.py > profiles/r02_ref_thread_sensitivity.jsonl
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_loader as rl, synth  # noqa: E402


def epe(a, b):
    return (a - b).pow(2).sum(1).sqrt().mean().item()


def run(model, x, threads):
    from spikingjelly.activation_based import functional
    torch.set_num_threads(threads)
    functional.reset_net(model)
    with torch.no_grad():
        return [f.clone() for f in model(x)["flow"]]


def per_layer_flips(model, x, t_a, t_b):
    """spike-flip rate of every neuron layer between the two thread counts (free running)."""
    from spikingjelly.activation_based import functional
    outs = {}

    def hook(name):
        def f(_m, _i, o):
            outs.setdefault(name, []).append(o.detach().clone())
        return f
    hs = [m.register_forward_hook(hook(n)) for n, m in model.named_modules() if type(m).__name__ == "Spiking_neuron"]
    for t in (t_a, t_b):
        torch.set_num_threads(t)
        functional.reset_net(model)
        with torch.no_grad():
            model(x)
    for h in hs:
        h.remove()
    rates = {n: (v[0] != v[1]).float().mean().item() for n, v in outs.items() if len(v) == 2}
    return rates


class gemms_in_fp64:
    """Context: every F.linear / F.conv2d / F.conv_transpose2d of the (unmodified) reference model computes in float64 and
    rounds the result to fp32 — the correctly rounded GEMM instead of the library's fp32 accumulation order.  Nothing else
    changes (BatchNorm, neurons, adds stay fp32).  This is the size of perturbation ANY re-implementation introduces."""

    def __enter__(self):
        import torch.nn.functional as F
        self.F, self.orig = F, (F.linear, F.conv2d, F.conv_transpose2d)
        lin, c2, ct = self.orig

        def d(t):
            return None if t is None else t.double()
        F.linear = lambda x, w, b=None: lin(x.double(), w.double(), d(b)).to(x.dtype)
        F.conv2d = lambda x, w, b=None, *a, **k: c2(x.double(), w.double(), d(b), *a, **k).to(x.dtype)
        F.conv_transpose2d = lambda x, w, b=None, *a, **k: ct(x.double(), w.double(), d(b), *a, **k).to(x.dtype)
        return self

    def __exit__(self, *a):
        self.F.linear, self.F.conv2d, self.F.conv_transpose2d = self.orig


def first_flip_and_rates(model, x, ctx_b):
    """per-layer spike-flip rates between the plain run and the run under ctx_b (free running)."""
    from spikingjelly.activation_based import functional
    import contextlib
    outs = {}

    def hook(name):
        def f(_m, _i, o):
            outs.setdefault(name, []).append(o.detach().clone())
        return f
    hs = [m.register_forward_hook(hook(n)) for n, m in model.named_modules() if type(m).__name__ == "Spiking_neuron"]
    flows = []
    for ctx in (contextlib.nullcontext(), ctx_b):
        functional.reset_net(model)
        with torch.no_grad(), ctx:
            flows.append(model(x)["flow"][-1].clone())
    for h in hs:
        h.remove()
    rates = {n: (v[0] != v[1]).float().mean().item() for n, v in outs.items() if len(v) == 2}
    return flows, rates


def main():
    n_threads = os.cpu_count() or 8
    cases = [("small_lif", dict(small="lif"), (2, 10, 96, 128)), ("small_psn", dict(small="psn"), (2, 10, 96, 128)),
             ("en4_lif_288x384", dict(neuron_type="lif", input_size=(288, 384)), (1, 10, 288, 384))]
    for name, kw, shape in cases:
        if "small" in kw:
            mc, sc = synth.small_config(kw["small"])
        else:
            mc, sc = rl.default_config(**kw)
        model = rl.build_reference_model(mc, sc, seed=0, train=False)
        model.load_state_dict(synth.synth_state_dict(model.state_dict(), 0), strict=True)
        x = synth.synth_voxels(*shape)
        f1 = run(model, x, 1)
        f1b = run(model, x, 1)
        fn = run(model, x, n_threads)
        f2 = run(model, x, 2)
        rates = per_layer_flips(model, x, 1, n_threads)
        first_nonzero = next((n for n, r in rates.items() if r > 0), None)
        out = {"case": name, "model": mc["name"], "threads": [1, n_threads],
               "flow_mag_px": f1[-1].pow(2).sum(1).sqrt().mean().item(),
               "epe_1_vs_1_rerun_px": epe(f1[-1], f1b[-1]),
               "epe_1_vs_2_threads_px": epe(f1[-1], f2[-1]),
               f"epe_1_vs_{n_threads}_threads_px": epe(f1[-1], fn[-1]),
               "max_abs_diff_px": (f1[-1] - fn[-1]).abs().max().item(),
               "neuron_layers": len(rates), "layers_with_flips": sum(r > 0 for r in rates.values()),
               "first_layer_with_a_flip": first_nonzero,
               "max_layer_flip_rate": max(rates.values()) if rates else None,
               "median_layer_flip_rate": sorted(rates.values())[len(rates) // 2] if rates else None}
        print(json.dumps(out), flush=True)
        # the reference against ITSELF with correctly rounded GEMMs (fp64 compute, fp32 result)
        torch.set_num_threads(n_threads)
        (fa, fb), r64 = first_flip_and_rates(model, x, gemms_in_fp64())
        order = list(r64.keys())
        first = next((n for n in order if r64[n] > 0), None)
        out = {"case": name, "perturbation": "GEMMs/convs computed in fp64 and rounded to fp32 (reference otherwise unmodified)",
               "flow_mag_px": fa.pow(2).sum(1).sqrt().mean().item(), "epe_px": epe(fa, fb),
               "max_abs_diff_px": (fa - fb).abs().max().item(), "neuron_layers": len(r64),
               "layers_with_flips": sum(v > 0 for v in r64.values()), "first_layer_with_a_flip": first,
               "flip_rate_of_that_layer": r64.get(first), "index_of_that_layer": order.index(first) if first else None,
               "max_layer_flip_rate": max(r64.values()), "median_layer_flip_rate": sorted(r64.values())[len(r64) // 2],
               "last_layer_flip_rate": r64[order[-1]]}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
