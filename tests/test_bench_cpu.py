"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line, and our arm
refuses to run without a GPU (there is no CPU fallback to time by accident)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "SDformerFlow fwd+bwd samples/s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_our_arm_needs_a_gpu():
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "GPU" in (r.stderr + r.stdout)


def test_committed_bench_line_keeps_the_contract():
    """profiles/r02_bench_n1.json (the default `python bench.py` line measured on the B200) carries every key of the bench
    contract, with the roofline of the dominant entry point consistent with its own numbers."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.load(open(os.path.join(root, "profiles", "r02_bench_n1.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["config"]["workload"] and "model" not in line["config"]
    assert line["vs_baseline"] is None and line["higher_is_better"] is True and line["n_gpus"] == 1 and line["warmup"] >= 3
    e2e, rf, cpu = line["e2e"], line["roofline"], line["cpu_baseline"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] <= line["value"] * 1.02
    assert rf["bound"] in ("hbm", "tensor") and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["traffic"] is None or 0.5 < rf["traffic"] / rf["algo_bytes_per_launch"] < 1.5     # no wasted re-reads
    assert cpu["kind"] in ("port", "reference") and cpu["cores"] >= 1 and cpu["sample"]
    assert line["gpu_launches"] > 0 and not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert abs(line["value"] - 4 * 1000.0 / line["ms_per_step"]) < 1e-6 * line["value"]
