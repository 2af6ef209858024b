#!/bin/bash
# Round-2 GPU regression pass on ONE GPU: GEMM harness -> parity tests -> peer loop-back tests -> bench (+ per-kernel table).
mkdir -p gpurun_out
echo "=== harness"; timeout 300 ./thinkdiff_mlre_b200/csrc/test_gemm.bin > gpurun_out/r02_harness.log 2>&1; echo "harness rc=$?"; grep -E "FAIL|RESULT|perf|KERNEL|rc=" gpurun_out/r02_harness.log | tail -30
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02_pytest.log
echo "=== bench"; timeout 600 python bench.py --steps ${1:-20} --warmup 5 --profile-out gpurun_out/r02_bench_profile.json > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "host_enqueue_ms_per_step", "step_tflops_per_gpu", "gpu_launches")})
    print("e2e", d.get("e2e")); print("clocks", d.get("clocks")); print("roofline", d.get("roofline"))
    tot = 0
    for k, v in d["kernels"].items():
        us = v["ms_per_launch"] * 1e3 * v["launches_per_step"]; tot += us
        print(k.ljust(22), f"{v['ms_per_launch']*1e3:7.1f} us x{v['launches_per_step']:.0f}", f"{v['achieved']:8.1f} {v['unit']}", f"frac {v['frac']:.3f}")
    print("sum of profiled kernels us/step", round(tot, 1))
except Exception as e:
    print("bench parse failed", e)
PY
for lib in thinkdiff_mlre_b200/libthinkdiff_b200_*.so; do  # A/B of alternative builds, if any were shipped
  [ -f "$lib" ] || continue
  tag=$(basename $lib .so | sed 's/libthinkdiff_b200_//')
  THINKDIFF_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps ${1:-20} --warmup 5 --no-cpu-baseline --no-e2e --no-eager-bar > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_$tag.json')); print('variant $tag', 'tok/s %.3fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], {t: round(v['ms_per_launch']*1e3) for t, v in d['kernels'].items()})"
done
