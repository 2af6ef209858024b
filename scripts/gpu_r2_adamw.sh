#!/bin/bash
# One GPU: A/B of the AdamW launch shape (TD_ADAMW) on the N=1 step, interleaved.
mkdir -p gpurun_out
A="--steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-eager-bar"
for rep in 1 2; do
  for mode in wide2 narrow8 wide8; do
    TD_ADAMW=$mode timeout 200 python bench.py $A > gpurun_out/adamw_${mode}_$rep.json 2> gpurun_out/adamw_${mode}_$rep.err
    python -c "
import json; d=json.load(open('gpurun_out/adamw_${mode}_$rep.json')); print('$mode'.ljust(8), 'tok/s %.3fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], {t: round(v['ms_per_launch']*1e3) for t, v in d['kernels'].items()})"
  done
done
