#!/bin/bash
# One 8-GPU box (8x the GPU-minutes -- keep it short): N=1 reference, then the peer step at N=8 with the folded protocol and the
# rank-rotated tile order switched off one at a time (developer env knobs), default configuration first (with e2e + timelines).
N=${1:-8}; STEPS=${2:-40}
mkdir -p gpurun_out
show() {
python -c "
import json,sys; d=json.load(open(sys.argv[2])); k=d.get('kernels',{}); print(sys.argv[1].ljust(18), 'N', d['n_gpus'], 'tok/s %.3fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], 'parity', (d.get('dp_parity') or {}).get('ok'), 'e2e %.2fM' % (((d.get('e2e') or {}).get('value') or 0)/1e6), {t: round(v['ms_per_launch']*1e3) for t, v in k.items()})" "$1" "$2" || tail -20 "${2%.json}.err"
}
run() {  # run TAG extra...
  tag=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps $STEPS --warmup 5 --no-cpu-baseline "$@" > gpurun_out/n8ab_$tag.json 2> gpurun_out/n8ab_$tag.err
  echo "rc=$?"; show $tag gpurun_out/n8ab_$tag.json
}
timeout 200 python bench.py --steps $STEPS --warmup 5 --no-cpu-baseline --no-eager-bar --no-e2e > gpurun_out/n8ab_n1.json 2> gpurun_out/n8ab_n1.err; show n1 gpurun_out/n8ab_n1.json
run default --timeline-out gpurun_out/n8ab_timeline_%r.json
TD_PEER_UNFOLDED=1 run unfolded --no-e2e --no-dp-parity
TD_PEER_NO_ROTATE=1 run norotate --no-e2e --no-dp-parity
python - $N <<'PY'
import json, sys
n = int(sys.argv[1])
try:
    a = json.load(open("gpurun_out/n8ab_n1.json"))
    for tag in ("default", "unfolded", "norotate"):
        b = json.load(open(f"gpurun_out/n8ab_{tag}.json"))
        print(f"{tag:10s} efficiency N={n}: {b['value'] / (n * a['value']):.3f}")
except Exception as e:
    print("efficiency table failed", e)
PY
