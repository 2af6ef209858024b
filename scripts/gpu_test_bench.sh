#!/bin/bash
# GPU regression pass: parity tests, smoke, bench (JSON to gpurun_out/bench.json + a compact per-kernel table)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps ${1:-50} --warmup 5 --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "host_enqueue_ms_per_step", "step_tflops_per_gpu", "step_frac_of_bf16_peak", "gpu_launches")})
print("e2e", d["e2e"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"]); print("roofline", d["roofline"])
tot = 0
for k, v in d["kernels"].items():
    us = v["ms_per_launch"] * 1e3 * v["launches_per_step"]; tot += us
    print(k.ljust(22), f"{v['ms_per_launch']*1e3:7.1f} us x{v['launches_per_step']:.0f}", f"{v['achieved']:8.1f} {v['unit']}", f"frac {v['frac']:.3f}", f"share {v['share_of_kernel_time']:.3f}")
print("sum of profiled kernels us/step", round(tot, 1))
PY
