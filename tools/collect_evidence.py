"""Copies the judged subset of gpurun_out/ (scratch) to profiles/ (tracked) and derives the ncu-only numbers bench.py
reports (profiles/<round>_ncu_numbers.json: DRAM traffic per launch of the dominant kernel, tensor-pipe utilisation of the
Q K^T V attention kernel).  Usage: python tools/collect_evidence.py r02"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def raw_rows(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names = rows[hdr]
    return names, [r for r in rows[hdr + 2:] if len(r) == len(names)]


def num(x):
    return float(x.replace(",", "")) if x not in ("", "n/a") else 0.0


def main():
    copied = []
    for f in sorted(os.listdir(SRC)):
        if f.startswith(R + "_") and f.endswith((".json", ".jsonl", ".txt")):
            shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))
            copied.append(f)
    raws = sorted(os.path.join(SRC, f) for f in os.listdir(SRC) if f.startswith(R + "_ncu_") and f.endswith(".raw.csv")
                  and "in_bench" not in f)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "summarise_ncu.py"), "full",
                           os.path.join(DST, f"{R}_ncu_full_kernels.md"), *raws])
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "summarise_ncu.py"), "list",
                           os.path.join(DST, f"{R}_launches_bench_step.md"), os.path.join(SRC, f"{R}_launches.csv"),
                           "python bench.py --steps 1 --warmup 3 --no-cpu-baseline --secondary off --graph off"])
    out = {"traffic": None, "traffic_source": None, "attn_tc_util": None}
    # dominant kernel's DRAM traffic: the 42 K2 launches of one training step, captured inside the bench
    out["traffic_by_entry"] = {}
    for kern, key, entry, must in (("lif_bwd", "lif_bwd_K2", "sdf_lif_bwd", "lif_bwd"), ("lif_fwd", "lif_fwd_K1", "sdf_lif_fwd", "lif_fwd"),
                                   ("bn_rows", "bn_bwd_apply", "sdf_bn_bwd_apply", "bn_rows_kernel<1>")):
        p = os.path.join(SRC, f"{R}_ncu_{kern}_in_bench.raw.csv")
        if not os.path.exists(p):
            continue
        names, rows = raw_rows(p)
        rows = [r for r in rows if must in r[names.index("Kernel Name")].replace(" ", "").replace("(int)", "")]
        if not rows:
            continue
        rd, wr, tm = names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum"), names.index("gpu__time_duration.sum")
        units = list(csv.reader(open(p)))[[i for i, r in enumerate(csv.reader(open(p))) if "Kernel Name" in r][0] + 1]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = sum(num(r[rd]) * scale.get(units[rd], 1.0) + num(r[wr]) * scale.get(units[wr], 1.0) for r in rows)
        per = tot / max(len(rows), 1)
        out[key] = {"launches": len(rows), "dram_bytes_per_launch": per, "dram_bytes_per_step": tot,
                    "sum_kernel_time_under_ncu": sum(num(r[tm]) for r in rows), "time_unit": units[tm]}
        out["traffic_by_entry"][entry] = {
            "dram_bytes_per_launch": per, "launches": len(rows),
            "source": (f"profiles/{R}_ncu_numbers.json <- ncu {'--set full' if 'sm__throughput.avg.pct_of_peak_sustained_elapsed' in names else '--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum'} -k regex:{kern} python bench.py --steps 1 --warmup 3 --graph off "
                       f"(dram__bytes_read.sum + dram__bytes_write.sum, mean over the {len(rows)} launches of {must} in one training step)")}
        if kern == "lif_bwd":
            out["traffic"] = per
            out["traffic_source"] = (f"profiles/{R}_ncu_numbers.json <- ncu --set full -k regex:lif_bwd -s 126 -c 42 python bench.py "
                                     "--steps 1 --warmup 3 --graph off (dram__bytes_read.sum + dram__bytes_write.sum, mean over the 42 "
                                     "K2 launches of one training step)")
    p = os.path.join(SRC, f"{R}_ncu_qktv2_fwd.raw.csv")
    if os.path.exists(p):
        names, rows = raw_rows(p)
        m = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
        kn, tm = names.index("Kernel Name"), names.index("gpu__time_duration.sum")
        per = [{"kernel": r[kn].split("(")[0], "pct_tensor_pipe_active": num(r[names.index(m)]), "us_under_ncu": num(r[tm])} for r in rows]
        out["attn_tc_util"] = {"pct_tensor_pipe_active": max(x["pct_tensor_pipe_active"] for x in per), "metric": m,
                               "per_launch": per, "workload": "(2,9,9) windows, M=10080 windows, nH=3 (<176,1> = shift-masked, <176,0> = unmasked)",
                               "source": f"profiles/{R}_ncu_full_kernels.md ({R}_ncu_qktv2_fwd)"}
    json.dump(out, open(os.path.join(DST, f"{R}_ncu_numbers.json"), "w"), indent=1)
    print("copied", copied)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
