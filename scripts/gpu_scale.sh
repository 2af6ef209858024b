#!/bin/bash
# multi-GPU bench sweep: usage gpu_scale.sh N [extra bench args]; prints compact lines. Env NCCL knobs are passed through.
N=$1; shift
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
  python - "$tag" gpurun_out/bench_n${N}_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read())
    print(sys.argv[1].ljust(14), "N", d["n_gpus"], "tok/s %.3fM" % (d["value"] / 1e6), "ms/step %.3f" % d["ms_per_step"], "TF/GPU %.0f" % d["step_tflops_per_gpu"],
          "host_ms %.2f" % d["host_enqueue_ms_per_step"], {k: round(v["ms_per_launch"] * 1e3) for k, v in d["kernels"].items() if k.startswith("gemm")}, d.get("allreduce_alone_ms"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
if [ "$N" = "1" ]; then
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/bench_n1_x.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_n1_x.json')); print('N1', 'tok/s %.3fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'])"
else
  run default "$@"
  NCCL_MAX_CTAS=8 run maxctas8 "$@"
  NCCL_MAX_CTAS=16 run maxctas16 "$@"
fi
