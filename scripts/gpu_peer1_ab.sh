#!/bin/bash
# One GPU: the peer data-parallel path at world size 1 next to the plain N=1 step; waits by stream memory operations vs the
# polling kernel. Isolates single-GPU effects of the peer step (SM sharing with the persistent GEMMs) from NVLink effects.
mkdir -p gpurun_out
summ() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read())
    print(sys.argv[1].ljust(14), "tok/s %.3fM" % (d["value"] / 1e6), "ms/step %.3f" % d["ms_per_step"], {t: round(v["ms_per_launch"] * 1e3) for t, v in d["kernels"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-1500:])
PY
}
A="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-eager-bar"
timeout 300 python bench.py $A > gpurun_out/p1_plain.json 2> gpurun_out/p1_plain.err; summ plain gpurun_out/p1_plain.json
timeout 300 python bench.py $A --dp peer > gpurun_out/p1_peer_memops.json 2> gpurun_out/p1_peer_memops.err; summ peer_memops gpurun_out/p1_peer_memops.json
TD_PEER_WAIT=kernel timeout 300 python bench.py $A --dp peer > gpurun_out/p1_peer_kernel.json 2> gpurun_out/p1_peer_kernel.err; summ peer_kernelwait gpurun_out/p1_peer_kernel.json
