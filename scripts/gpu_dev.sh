#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-eager-bar --timeline-out gpurun_out/n1_timeline.json > gpurun_out/n1_tl.json 2> gpurun_out/n1_tl.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/n1_tl.json')); print('tok/s %.3fM ms/step %.3f' % (d['value']/1e6, d['ms_per_step']))
t=json.load(open('gpurun_out/n1_timeline.json')); L=t['launches']
packs=[i for i,l in enumerate(L) if l['tag']=='pack_varlen']
a,b=packs[1],packs[2]
for l in sorted(L[a:b+1], key=lambda x:x['t0']):
    print(f"  {l['tag']:22s} s{l['stream']} {l['t0']:8.3f} -> {l['t1']:8.3f} ({(l['t1']-l['t0'])*1e3:5.0f} us)")
PY
