#!/bin/bash
# run each GEMM variant in its own process (a device trap kills the context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gemm_test.log 2>&1
for v in 0 1 2 3 4 5; do
  echo "=== variant $v" >> gpurun_out/gemm_test.log
  timeout 120 ./thinkdiff_mlre_b200/csrc/test_gemm.bin ${1:-full} $v >> gpurun_out/gemm_test.log 2>&1
  echo "exit=$?" >> gpurun_out/gemm_test.log
done
cat gpurun_out/gemm_test.log
