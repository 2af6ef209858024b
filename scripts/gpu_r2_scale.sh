#!/bin/bash
# Round-2 scaling pass on one multi-GPU box: N=1, then the peer data-parallel step at every N given, then NCCL-sharded at the
# largest N for comparison. Usage: gpu_r2_scale.sh "2 4 8" [steps]
NS=${1:-"2 4 8"}; STEPS=${2:-20}
mkdir -p gpurun_out
summ() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read())
    k = d.get("kernels", {})
    print(sys.argv[1].ljust(14), "N", d["n_gpus"], "tok/s %.3fM" % (d["value"] / 1e6), "ms/step %.3f" % d["ms_per_step"], "host_ms %.2f" % d.get("host_enqueue_ms_per_step", 0),
          "dp", d["config"].get("dp_exchange"), "parity", (d.get("dp_parity") or {}).get("ok"), "e2e %.2fM" % (((d.get("e2e") or {}).get("value") or 0) / 1e6),
          {t: round(v["ms_per_launch"] * 1e3) for t, v in k.items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-2500:])
PY
}
run() {  # run N dp extra...
  n=$1; dp=$2; shift 2
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $n --steps $STEPS --warmup 5 --no-cpu-baseline --dp $dp "$@" > gpurun_out/r02s_n${n}_$dp.json 2> gpurun_out/r02s_n${n}_$dp.err
  echo "rc=$?"; summ n${n}_$dp gpurun_out/r02s_n${n}_$dp.json
}
timeout 300 python bench.py --steps $STEPS --warmup 5 --no-cpu-baseline --no-eager-bar > gpurun_out/r02s_n1.json 2> gpurun_out/r02s_n1.err; summ n1 gpurun_out/r02s_n1.json
last=1
for n in $NS; do
  run $n peer --timeline-out gpurun_out/r02s_n${n}_timeline_%r.json
  last=$n
done
[ -n "$WITH_SHARDED" ] && run $last sharded --no-e2e
python - $NS <<'PY'
import json, sys
try:
    v1 = json.load(open("gpurun_out/r02s_n1.json"))["value"]
    for n in sys.argv[1:]:
        d = json.load(open(f"gpurun_out/r02s_n{n}_peer.json"))
        print(f"efficiency N={n}: {d['value'] / (int(n) * v1):.3f}   e2e eff {((d.get('e2e') or {}).get('value') or 0) / (int(n) * json.load(open('gpurun_out/r02s_n1.json'))['e2e']['value']):.3f}")
except Exception as e:
    print("efficiency table failed", e)
PY
