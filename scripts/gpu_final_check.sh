#!/bin/bash
# What the driver runs at round end, on one GPU: the whole GPU suite, smoke(), the default bench line and the reference arm.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py --steps ${1:-20} --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/final_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/final_bench.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "step_frac_of_bf16_burst", "host_enqueue_ms_per_step")})
print("e2e", d.get("e2e")); print("clocks", d.get("clocks")); print("roofline", d.get("roofline")); print("cpu", d.get("cpu_baseline")); print("eager", d.get("eager_bar"))
for k, v in d["kernels"].items():
    print(k.ljust(22), f"{v['ms_per_launch']*1e3:7.1f} us x{v['launches_per_step']:.0f}", f"{v['achieved']:8.1f} {v['unit']}", f"frac {v['frac']:.3f}")
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm rc=$?"; cat gpurun_out/final_bench_reference.json | cut -c1-400
