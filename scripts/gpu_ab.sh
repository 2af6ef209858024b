#!/bin/bash
# One GPU: parity tests with the default library and with $TEST_LIB, N=1 timeline with $TEST_LIB, interleaved A/B of all builds.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_aligner.py -m gpu -q -x > gpurun_out/ab3_pytest_default.log 2>&1; echo "pytest(default) rc=$?"; tail -2 gpurun_out/ab3_pytest_default.log
if [ -n "$TEST_LIB" ]; then
  export THINKDIFF_B200_LIB=$PWD/thinkdiff_mlre_b200/libthinkdiff_b200_$TEST_LIB.so
  timeout 600 python -m pytest tests/test_gpu_aligner.py tests/test_gpu_gemm.py tests/test_gpu_peer.py tests/test_gpu_optim.py -m gpu -q -x > gpurun_out/ab3_pytest.log 2>&1; echo "pytest($TEST_LIB) rc=$?"; tail -2 gpurun_out/ab3_pytest.log
  bash scripts/gpu_dev.sh
  unset THINKDIFF_B200_LIB
fi
A="--steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-eager-bar"
for rep in 1 2 3; do
  timeout 300 python bench.py $A > gpurun_out/ab3_default_$rep.json 2> gpurun_out/ab3_default_$rep.err
  python -c "
import json; d=json.load(open('gpurun_out/ab3_default_$rep.json')); print('default  ', 'tok/s %.3fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], {t: round(v['ms_per_launch']*1e3) for t, v in d['kernels'].items()})"
  for lib in thinkdiff_mlre_b200/libthinkdiff_b200_*.so; do
    [ -f "$lib" ] || continue
    tag=$(basename $lib .so | sed 's/libthinkdiff_b200_//')
    THINKDIFF_B200_LIB=$PWD/$lib timeout 300 python bench.py $A > gpurun_out/ab3_${tag}_$rep.json 2> gpurun_out/ab3_${tag}_$rep.err
    python -c "
import json; d=json.load(open('gpurun_out/ab3_${tag}_$rep.json')); print('variant $tag', 'tok/s %.3fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], {t: round(v['ms_per_launch']*1e3) for t, v in d['kernels'].items()})"
  done
done
