#!/bin/bash
# ncu evidence for the bench step: (1) launch list with device times, (2) one --set full capture of the GEMM launches
# of one step (with source). Run under gpurun; results land in gpurun_out/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-pipeline"  # sequential step: same kernels, one stream, fixed launch order
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
K=${1:-gemm_bf16_kernel}
SKIP=${2:-10}
COUNT=${3:-5}
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $COUNT -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
