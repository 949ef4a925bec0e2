"""Turns `ncu --page raw --csv` dumps (gpurun_out/*.raw.csv) and the launch-list CSV into the markdown kept under profiles/.
Usage: python tools/summarise_ncu.py full  OUT.md  a.raw.csv b.raw.csv ...
       python tools/summarise_ncu.py list  OUT.md  launches.csv  "<command line that was profiled>" """
import csv
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_tmem.sum", "smsp__inst_executed.sum",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]


def full(out, files):
    with open(out, "w") as f:
        f.write("# `ncu --set full --clock-control none --import-source on` (tools/ncu_targets.py at BASELINE sizes; raw page)\n")
        for path in files:
            rows = list(csv.reader(open(path)))
            hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
            names, units = rows[hdr], rows[hdr + 1]
            kcol = names.index("Kernel Name")
            for r in rows[hdr + 2:]:
                if len(r) != len(names):
                    continue
                f.write(f"\n## {path.split('/')[-1].replace('.raw.csv', '')}: `{r[kcol]}`\n\n| metric | value | unit |\n|---|---|---|\n")
                for m in KEEP:
                    if m in names:
                        i = names.index(m)
                        f.write(f"| {m} | {r[i]} | {units[i]} |\n")
                # every tensor-pipe metric the tool knows, whatever it is called on this chip
                for i, m in enumerate(names):
                    if (("pipe_tensor" in m or "mem_tensor" in m or "pipe_tmem" in m) and ".avg.pct_of_peak_sustained_active" in m
                            and m not in KEEP and r[i] not in ("0", "", "n/a")):
                        f.write(f"| {m} | {r[i]} | {units[i]} |\n")


def launch_list(out, path, cmd):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    tot = defaultdict(lambda: [0.0, 0])
    for r in rows:
        name = r[4].split("(")[0].split("<")[0]
        tot[name][0] += float(r[-1].replace(",", "")) / 1e6
        tot[name][1] += 1
    total = sum(v[0] for v in tot.values())
    ours = sum(v[0] for k, v in tot.items() if "sdf::" in k or k.startswith("void sdf") or "sdf" in k.split()[-1])
    with open(out, "w") as f:
        f.write(f"# ncu launch list: `{cmd}`\n\n{len(rows)} launches. Cold-cache, serialised durations: compare SHARES, not absolutes.\n\n")
        f.write(f"* libsdf_b200 kernels: **{100 * ours / total:.1f}%** of summed kernel time\n\n| share | total ms | launches | kernel |\n|---|---|---|---|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:45]:
            f.write(f"| {100 * v[0] / total:.2f}% | {v[0]:.2f} | {v[1]} | `{k[:90]}` |\n")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3:])
    else:
        launch_list(sys.argv[2], sys.argv[3], sys.argv[4])
