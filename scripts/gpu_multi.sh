#!/bin/bash
# Multi-GPU check of a GEMM / peer change, usage: gpu_multi.sh N. (1) GEMM / peer / aligner / data-parallel parity tests,
# (2) N=1 A/B of the default library against every libthinkdiff_b200_<tag>.so (N=1 only: alternative builds may have an older
# ABI for the multi-GPU entries), (3) the peer step at N with per-rank timelines.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_peer.py tests/test_gpu_aligner.py tests/test_gpu_dp.py tests/test_gpu_optim.py -m gpu -q -x > gpurun_out/m3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/m3_pytest.log
A="--steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-eager-bar"
show() {
python -c "
import json,sys; d=json.load(open(sys.argv[2])); print(sys.argv[1].ljust(14), 'N', d['n_gpus'], 'tok/s %.3fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], 'parity', (d.get('dp_parity') or {}).get('ok'), {t: round(v['ms_per_launch']*1e3) for t, v in d['kernels'].items()})" "$1" "$2" || tail -20 "${2%.json}.err"
}
for rep in 1 2 3; do
  timeout 300 python bench.py $A > gpurun_out/m3_default_$rep.json 2> gpurun_out/m3_default_$rep.err; show default gpurun_out/m3_default_$rep.json
  for lib in thinkdiff_mlre_b200/libthinkdiff_b200_*.so; do
    [ -f "$lib" ] || continue
    tag=$(basename $lib .so | sed 's/libthinkdiff_b200_//')
    THINKDIFF_B200_LIB=$PWD/$lib timeout 300 python bench.py $A > gpurun_out/m3_${tag}_$rep.json 2> gpurun_out/m3_${tag}_$rep.err; show "variant_$tag" gpurun_out/m3_${tag}_$rep.json
  done
done
for rep in 1 2; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --timeline-out gpurun_out/m3_n${N}_timeline_%r.json > gpurun_out/m3_n${N}_peer_$rep.json 2> gpurun_out/m3_n${N}_peer_$rep.err
  echo "rc=$?"; show n${N}_peer gpurun_out/m3_n${N}_peer_$rep.json
done
