#!/bin/bash
# Light refresh of the round evidence after late changes: tests, smoke, bench lines, microbench (no ncu captures).
set -x
R=r02
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4 > gpurun_out/${R}_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -5 > gpurun_out/${R}_smoke.txt
python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${R}_bench_n1.json
python bench.py --workload infer --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${R}_bench_infer_n1.json
python tools/microbench.py 2>/dev/null > gpurun_out/${R}_microbench.jsonl
cat gpurun_out/${R}_pytest_gpu.txt gpurun_out/${R}_smoke.txt
cut -c1-400 gpurun_out/${R}_bench_n1.json
