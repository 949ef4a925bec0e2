#!/bin/bash
# Round evidence run (one B200): tests, smoke, bench (both arms), microbench, ncu launch list + full captures.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4 > gpurun_out/r01_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/r01_smoke.txt
python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r01_bench_n1.json
python bench.py --workload infer --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r01_bench_infer_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r01_bench_reference_arm.json
python tools/microbench.py 2>/dev/null > gpurun_out/r01_microbench.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r01_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lif_fwd_kernel -s 126 -c 42 -o gpurun_out/r01_ncu_lif_fwd_in_bench python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lif_bwd_kernel -s 1 -c 1 -o gpurun_out/r01_ncu_lif_bwd_final python tools/ncu_targets.py lif_bwd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qkgate_kernel -s 1 -c 1 -o gpurun_out/r01_ncu_qkgate_final python tools/ncu_targets.py qkgate > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qktv_kernel -s 1 -c 1 -o gpurun_out/r01_ncu_qktv_final python tools/ncu_targets.py qktv > /dev/null 2>&1
ls -la gpurun_out/ | tail -20
cat gpurun_out/r01_pytest_gpu.txt gpurun_out/r01_smoke.txt
cut -c1-600 gpurun_out/r01_bench_n1.json
cut -c1-300 gpurun_out/r01_bench_reference_arm.json
