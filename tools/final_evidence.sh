#!/bin/bash
# Round-2 evidence run (one B200): tests, smoke, bench (both arms + secondary configs), microbenchmarks, ncu launch list and
# full captures.  Everything lands in gpurun_out/ (scratch); tools/collect_evidence.py copies the judged subset to profiles/.
set -x
R=r02
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4 > gpurun_out/${R}_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -5 > gpurun_out/${R}_smoke.txt
python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${R}_bench_n1.json
python bench.py --workload infer --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${R}_bench_infer_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${R}_bench_reference_arm.json
python tools/microbench.py 2>/dev/null > gpurun_out/${R}_microbench.jsonl
python tools/bench_gemm.py 2>/dev/null > gpurun_out/${R}_bench_gemm.jsonl
python tools/bench_small_conv.py 2>/dev/null > gpurun_out/${R}_bench_small_conv.txt
python tools/bench_conv_bwd.py 2>/dev/null > gpurun_out/${R}_bench_conv_bwd.jsonl
python tools/profile_step.py lif --aten > gpurun_out/${R}_profile_step.txt 2>&1
for m in 0 1 2; do SDF_WGRAD_DEBUG=$m python tools/bench_wgrad_dbg.py 2>/dev/null | grep case; done > gpurun_out/${R}_bench_wgrad_modes.jsonl
./tools/ubench/tma_stream > gpurun_out/${R}_ubench_tma_stream.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --secondary off --graph off > gpurun_out/${R}_ncu_bench.log 2>&1
# full captures: keep the raw-page CSV (the .ncu-rep files exceed the 64 MiB return limit)
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
# in-bench captures (every launch of a kernel family over one training step): DRAM bytes + duration only — the full set on
# 42-103 launches costs ~6 GPU-minutes per family (the first r02 run of this script hit its 25-minute limit there)
cap_light() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$rx -s $skip -c $cnt -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
BENCH1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --secondary off --graph off"
cap_light ${R}_ncu_lif_bwd_in_bench lif_bwd 126 42 $BENCH1
cap_light ${R}_ncu_bn_rows_in_bench bn_rows_kernel 309 103 $BENCH1
cap_light ${R}_ncu_lif_fwd_in_bench lif_fwd_kernel 126 42 $BENCH1
cap ${R}_ncu_lif_fwd lif_fwd_kernel 2 1 python tools/ncu_targets.py lif_fwd
cap ${R}_ncu_lif_bwd lif_bwd 1 1 python tools/ncu_targets.py lif_bwd
cap ${R}_ncu_qktv2_fwd qktv2_kernel 1 2 python tools/ncu_targets.py qktv
cap ${R}_ncu_qktv2_bwd qktv2_bwd_kernel 0 3 python tools/ncu_targets.py qktv
cap ${R}_ncu_lin_fwd gemm_kernel 2 1 python tools/ncu_gemm_targets.py lin_fwd
cap ${R}_ncu_conv_fwd gemm_kernel 2 1 python tools/ncu_gemm_targets.py conv_fwd
cap ${R}_ncu_lin_dgrad gemm_kernel 2 1 python tools/ncu_gemm_targets.py lin_dgrad
cap ${R}_ncu_conv_dgrad gemm_kernel 2 1 python tools/ncu_gemm_targets.py conv_dgrad
cap ${R}_ncu_lin_wgrad wgrad_kernel 2 1 python tools/ncu_gemm_targets.py lin_wgrad
cap ${R}_ncu_conv_wgrad wgrad_kernel 2 1 python tools/ncu_gemm_targets.py conv_wgrad
cap ${R}_ncu_deconv_fwd gemm_kernel 8 4 python tools/ncu_gemm_targets.py deconv_fwd
ls -la gpurun_out/ | grep ${R}_ | tail -40
cat gpurun_out/${R}_pytest_gpu.txt gpurun_out/${R}_smoke.txt
cut -c1-700 gpurun_out/${R}_bench_n1.json
cut -c1-400 gpurun_out/${R}_bench_reference_arm.json
