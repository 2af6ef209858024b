#!/bin/bash
# Bring-up of the experimental peer-memory data-parallel path. Usage:
#   gpurun --timeout 600 -- 'bash scripts/gpu_peer.sh 1'                  # one GPU: loop-back tests of every kernel
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_peer.sh 2'          # two GPUs: IPC mapping, training parity, A/B bench
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_peer.sh 8 bench'    # A/B bench only
# Every step runs under its own timeout; a peer wait that makes no progress traps after 30 s instead of hanging.
N=${1:-1}; MODE=${2:-all}
mkdir -p gpurun_out
if [ "$MODE" != "bench" ]; then
  TD_TEST_PEER=1 timeout 600 python -m pytest tests/test_gpu_peer.py -q -x 2>&1 | tail -15
fi
if [ "$N" -gt 1 ]; then
  for tag in nccl peer; do
    extra=""; [ "$tag" = "peer" ] && extra="--peer"
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-e2e $extra > gpurun_out/peer_ab_n${N}_$tag.json 2> gpurun_out/peer_ab_n${N}_$tag.err
    python - "$tag" gpurun_out/peer_ab_n${N}_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read())
    print(sys.argv[1].ljust(6), "N", d["n_gpus"], "tok/s %.3fM" % (d["value"] / 1e6), "ms/step %.3f" % d["ms_per_step"], "TF/GPU %.0f" % d["step_tflops_per_gpu"],
          "loss", d.get("final_loss"), {k: round(v["ms_per_launch"] * 1e3) for k, v in d["kernels"].items() if k.startswith(("gemm", "adamw"))})
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-1500:])
PY
  done
fi
