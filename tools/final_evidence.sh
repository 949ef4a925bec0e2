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
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph off > gpurun_out/r01_ncu_bench.log 2>&1
# full captures: keep only the raw-page CSV (the .ncu-rep files exceed the 64 MiB return limit)
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
cap r01_ncu_lif_fwd_in_bench lif_fwd_kernel 126 42 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph off
cap r01_ncu_lif_fwd_final lif_fwd_kernel 2 2 python tools/ncu_targets.py lif_fwd
cap r01_ncu_lif_bwd_final lif_bwd_kernel 1 1 python tools/ncu_targets.py lif_bwd
cap r01_ncu_qkgate_final qkgate_kernel 1 1 python tools/ncu_targets.py qkgate
cap r01_ncu_qktv2_fwd qktv2_kernel 1 2 python tools/ncu_targets.py qktv
cap r01_ncu_qktv2_bwd qktv2_bwd_kernel 0 3 python tools/ncu_targets.py qktv
cap r01_ncu_qktv_v1_large qktv_kernel 1 1 python tools/ncu_targets.py qktv
./tools/ubench/mma_chain > gpurun_out/r01_ubench_mma_chain.jsonl 2>&1
ls -la gpurun_out/ | tail -20
cat gpurun_out/r01_pytest_gpu.txt gpurun_out/r01_smoke.txt
cut -c1-600 gpurun_out/r01_bench_n1.json
cut -c1-300 gpurun_out/r01_bench_reference_arm.json
