#!/bin/bash
# Round-2 scaling pass on one 8-GPU box (8x the GPU-minutes: keep it short). N=1 on the same box, then the default data-parallel
# step (peer) at N with e2e + per-rank timelines, then BASELINE configs 5 and 4 at N. Usage: gpu_r2_n8.sh [N] [steps]
N=${1:-8}; STEPS=${2:-20}
mkdir -p gpurun_out
summ() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read())
    k = d.get("kernels", {})
    print(sys.argv[1].ljust(14), "N", d["n_gpus"], "tok/s %.3fM" % (d["value"] / 1e6), "ms/step %.3f" % d["ms_per_step"], "host_ms %.2f" % d.get("host_enqueue_ms_per_step", 0),
          "dp", d["config"].get("dp_exchange"), "parity", (d.get("dp_parity") or {}).get("ok"), "e2e %.2fM" % (((d.get("e2e") or {}).get("value") or 0) / 1e6),
          "h2d GB/s/rank %.1f" % ((d.get("e2e") or {}).get("h2d_gbs_per_rank") or 0), "clk", (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"),
          {t: round(v["ms_per_launch"] * 1e3) for t, v in k.items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-2500:])
PY
}
run() {  # run TAG N extra...
  tag=$1; n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $n --steps $STEPS --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r02s_$tag.json 2> gpurun_out/r02s_$tag.err
  echo "rc=$?"; summ $tag gpurun_out/r02s_$tag.json
}
timeout 200 python bench.py --steps $STEPS --warmup 5 --no-cpu-baseline --no-eager-bar > gpurun_out/r02s_n1.json 2> gpurun_out/r02s_n1.err; summ n1 gpurun_out/r02s_n1.json
run n${N}_cfg2 $N --timeline-out gpurun_out/r02s_n${N}_timeline_%r.json
run n${N}_cfg5 $N --workload cfg5 --no-e2e
run n${N}_cfg4 $N --workload cfg4
python - $N <<'PY'
import json, sys
n = int(sys.argv[1])
try:
    a, b = json.load(open("gpurun_out/r02s_n1.json")), json.load(open(f"gpurun_out/r02s_n{n}_cfg2.json"))
    print(f"cfg2 efficiency N={n}: {b['value'] / (n * a['value']):.3f}   e2e efficiency {b['e2e']['value'] / (n * a['e2e']['value']):.3f}")
    b = json.load(open(f"gpurun_out/r02s_n{n}_cfg5.json"))
    print(f"cfg5 N={n}: {b['value'] / 1e6:.2f} M tok/s, step TFLOP/s/GPU {b['step_tflops_per_gpu']:.0f}, frac burst {b['step_frac_of_bf16_burst']:.3f}")
except Exception as e:
    print("efficiency table failed", e)
PY
