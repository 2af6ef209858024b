#!/bin/bash
# One GPU: parity tests of the GEMM / aligner paths, then the N=1 bench with the default library and with every alternative
# build shipped as thinkdiff_mlre_b200/libthinkdiff_b200_<tag>.so (A/B on the same box, interleaved twice).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_aligner.py tests/test_gpu_optim.py -m gpu -q -x > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ab_pytest.log
A="--steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-eager-bar"
for rep in 1 2; do
  timeout 300 python bench.py $A > gpurun_out/ab_new_$rep.json 2> gpurun_out/ab_new_$rep.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_new_$rep.json')); print('default  ', 'tok/s %.3fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], {t: round(v['ms_per_launch']*1e3) for t, v in d['kernels'].items()})"
  for lib in thinkdiff_mlre_b200/libthinkdiff_b200_*.so; do
    [ -f "$lib" ] || continue
    tag=$(basename $lib .so | sed 's/libthinkdiff_b200_//')
    THINKDIFF_B200_LIB=$PWD/$lib timeout 300 python bench.py $A > gpurun_out/ab_${tag}_$rep.json 2> gpurun_out/ab_${tag}_$rep.err
    python -c "
import json; d=json.load(open('gpurun_out/ab_${tag}_$rep.json')); print('variant $tag', 'tok/s %.3fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], {t: round(v['ms_per_launch']*1e3) for t, v in d['kernels'].items()})"
  done
done
