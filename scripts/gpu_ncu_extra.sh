#!/bin/bash
# ncu --set full of the kernels the N=1 train-step capture (gpu_ncu.sh) does not contain: the scattered weight-gradient GEMM,
# the slot-summing AdamW and the finisher of the peer data-parallel step (world size 1: same kernels, local "peers"), and the
# lm_head GEMMs + cross-entropy pass of the frozen T5 head.
mkdir -p gpurun_out
CMD1="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-eager-bar --dp peer"
$CMD1 > gpurun_out/ncu_extra_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel<2, true, true, 5>|gemm_bf16_kernelILi2ELb1ELb1ELi5|adamw_slots|finish_kernel" -s 12 -c 6 -f -o gpurun_out/r02_prof_peer $CMD1 > gpurun_out/ncu_extra1.log 2>&1
echo "peer capture rc=$?"; tail -2 gpurun_out/ncu_extra1.log
CMD2="python scripts/bench_lm_head.py 8192"
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel|masked_ce" -s 6 -c 3 -f -o gpurun_out/r02_prof_lm_head $CMD2 > gpurun_out/ncu_extra2.log 2>&1
echo "lm_head capture rc=$?"; tail -2 gpurun_out/ncu_extra2.log
ls -la gpurun_out/*.ncu-rep
