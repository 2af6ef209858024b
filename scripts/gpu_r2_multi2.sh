#!/bin/bash
# Round-2 multi-GPU pass, usage: gpu_r2_multi2.sh N. (1) the whole GPU suite (multi-GPU tests included), (2) N=1 bench on the
# same box, (3) bench at N with every gradient-exchange mode (peer with a per-rank timeline).
N=${1:-2}
mkdir -p gpurun_out
summ() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read())
    k = d.get("kernels", {})
    print(sys.argv[1].ljust(16), "N", d["n_gpus"], "tok/s %.3fM" % (d["value"] / 1e6), "ms/step %.3f" % d["ms_per_step"], "host_ms %.2f" % d.get("host_enqueue_ms_per_step", 0),
          "dp", d["config"].get("dp_exchange"), "parity", (d.get("dp_parity") or {}).get("ok"), "e2e", round(((d.get("e2e") or {}).get("value") or 0) / 1e6, 2),
          {t: round(v["ms_per_launch"] * 1e3) for t, v in k.items()})
    if d.get("dp_parity"): print("   dp_parity", d["dp_parity"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-2500:])
PY
}
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02m_pytest.log
echo "=== N=1 on this box"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-eager-bar > gpurun_out/r02m_n1.json 2> gpurun_out/r02m_n1.err; summ n1_default gpurun_out/r02m_n1.json
echo "=== bench N=$N"
for dp in peer sharded allreduce; do
  extra="--no-e2e"; [ "$dp" = "peer" ] && extra="--timeline-out gpurun_out/r02m_n${N}_timeline_%r.json"
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --dp $dp $extra > gpurun_out/r02m_n${N}_$dp.json 2> gpurun_out/r02m_n${N}_$dp.err
  echo "rc=$?"; summ n${N}_$dp gpurun_out/r02m_n${N}_$dp.json
done
