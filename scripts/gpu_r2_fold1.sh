#!/bin/bash
# One GPU: the folded peer protocol in loop-back (tests), then the peer step at world size 1 folded / unfolded next to the plain step.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_peer.py tests/test_gpu_gemm.py tests/test_gpu_aligner.py -m gpu -q -x > gpurun_out/f1_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f1_pytest.log
show() {
python -c "
import json,sys; d=json.load(open(sys.argv[2])); print(sys.argv[1].ljust(16), 'tok/s %.3fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], 'launches/step', d['gpu_launches']/d['steps'], {t: round(v['ms_per_launch']*1e3) for t, v in d['kernels'].items()})" "$1" "$2" || tail -20 "${2%.json}.err"
}
A="--steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-eager-bar"
timeout 200 python bench.py $A > gpurun_out/f1_plain.json 2> gpurun_out/f1_plain.err; show plain gpurun_out/f1_plain.json
timeout 200 python bench.py $A --dp peer > gpurun_out/f1_peer_folded.json 2> gpurun_out/f1_peer_folded.err; show peer_folded gpurun_out/f1_peer_folded.json
TD_PEER_UNFOLDED=1 timeout 200 python bench.py $A --dp peer > gpurun_out/f1_peer_unfolded.json 2> gpurun_out/f1_peer_unfolded.err; show peer_unfolded gpurun_out/f1_peer_unfolded.json
