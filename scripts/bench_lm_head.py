#!/usr/bin/env python
"""Timing of SURVEY section 8 f-1's first slice on one B200: frozen T5-XXL output head (4096 -> 32128) + CE + backward to the
decoder output, R = B * T rows (default 64 x 128, configs/train_thinkdiff_lvlm_ccsbu.yaml batch 32/GPU x max_txt_len 128 is half of
it), next to the same expression in eager PyTorch under bf16 autocast (cuBLASLt + ATen) on the same GPU.

    python scripts/bench_lm_head.py [R] > profiles/rNN_lm_head.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import thinkdiff_mlre_b200 as td
from thinkdiff_mlre_b200 import _lib as L

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K, V, reps = 4096, 32128, 10
gen = torch.Generator(device="cuda").manual_seed(0)
W = (torch.randn((V, K), generator=gen, device="cuda") * 0.02).to(torch.bfloat16)
seqs = [torch.randn((R, K), generator=gen, device="cuda").to(torch.bfloat16) for _ in range(2)]
labels = torch.randint(0, V, (R,), generator=gen, device="cuda")
labels[torch.rand((R,), generator=gen, device="cuda") < 0.3] = -100


def timed(fn):
    for i in range(3):
        fn(seqs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(seqs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ours = timed(lambda s: td.ops.lm_head_ce(s, W, labels))
L.profile_enable(True)
for i in range(reps):
    td.ops.lm_head_ce(seqs[i % 2], W, labels)
torch.cuda.synchronize()
prof = L.profile_report()
L.profile_enable(False)
Wf = W.float().requires_grad_(False)


def eager(s):
    s = s.detach().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = torch.nn.functional.linear(s, Wf)
        loss = torch.nn.CrossEntropyLoss(ignore_index=-100)(logits.view(-1, V), labels)
    loss.backward()
    return s.grad


ref = timed(eager)
flops = 2 * 2.0 * R * K * V
out = {"rows": R, "d_model": K, "vocab": V, "ms": ours, "tflops": flops / ours / 1e9, "eager_ms": ref, "speedup_vs_eager": ref / ours,
       "logits_bytes": 2 * R * V, "kernels": {t: {"ms_per_launch": r["ms"] / r["launches"], "rate": r["work"] / (r["ms"] * 1e-3)} for t, r in prof.items()},
       "note": "loss + gradient w.r.t. the decoder output; logits bf16 materialised once, their gradient written in place; eager = F.linear + CrossEntropyLoss under bf16 autocast with the fp32 frozen weight (autocast casts it per call)"}
print(json.dumps(out))
