#!/bin/bash
mkdir -p gpurun_out
python scripts/dev/debug_scaler.py 2>&1 | tail -20
python scripts/dev/debug_scaler.py noscaler 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_optim.py tests/test_gpu_pack.py -q 2>&1 | tail -15
