#!/bin/bash
# ncu evidence for the train step's kernels (one GPU): launch list with device times + one --set full capture with source.
mkdir -p gpurun_out
CMD="python scripts/dev/one_step.py 4"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel|norm_mse_bwd|pack_rows|adamw256" -s 18 -c 9 -f -o gpurun_out/r02_prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
