#!/usr/bin/env python
"""The kernel-level bar on the same box: the reference's aligner as the reference runs it -- an eager
``nn.Sequential(Linear, GELU, Linear, T5LayerNorm)`` under bf16 autocast on the ZERO-PADDED ``[B, L_max, Din]`` batch
(thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:58-63, :585; collater padding ...mllama_embed_2.py:124-127), masked MSE on
the valid rows, ``backward()``, torch fused AdamW (runners/runner_base.py:122-127) -- timed on one B200 with the same synthetic
config-2 batches and the same token accounting (valid rows only) as ``bench.py``. cuBLASLt GEMMs + unfused ATen kernels: no
kernel of this repo runs here. Prints one JSON line; a comparator, not a product path.

    python scripts/bench_eager_bar.py [--steps 50] [--warmup 5] [--packed]      (--packed: same module on the packed rows)
"""
import argparse
import json
import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DIN, D, SEQS, MAX_LEN, NUM_BATCHES = 3584, 4096, 64, 256, 4


class T5Norm(nn.Module):  # transformers' T5LayerNorm.forward, restated (five lines)
    def __init__(self, d, eps=1e-6):
        super().__init__()
        self.weight, self.variance_epsilon = nn.Parameter(torch.ones(d)), eps

    def forward(self, h):
        var = h.to(torch.float32).pow(2).mean(-1, keepdim=True)
        h = h * torch.rsqrt(var + self.variance_epsilon)
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            h = h.to(self.weight.dtype)
        return self.weight * h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--packed", action="store_true", help="feed the packed [M, Din] rows instead of the padded batch")
    args = ap.parse_args()
    from thinkdiff_mlre_b200.train_step import make_reference_optimizer, synthetic_lvlm_batch

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    try:
        from transformers.models.t5.modeling_t5 import T5LayerNorm as Norm
    except Exception:
        Norm = T5Norm
    model = nn.Sequential(nn.Linear(DIN, D), nn.GELU(), nn.Linear(D, D), Norm(D)).to(dev)
    opt = make_reference_optimizer(model)

    batches = []
    for j in range(NUM_BATCHES):
        b = synthetic_lvlm_batch(SEQS, MAX_LEN, DIN, D, seed=1234 + 1 + 1000 * j, pin=False)
        lens, l_max = b.lens.tolist(), int(b.lens.max())
        x = torch.zeros((SEQS, l_max, DIN), dtype=torch.bfloat16)
        t = torch.zeros((SEQS, l_max, D), dtype=torch.bfloat16)
        mask = torch.zeros((SEQS, l_max), dtype=torch.bool)
        for i, (s, n) in enumerate(zip(b.src_row_start.tolist(), lens)):  # the reference collater's pad / stack / mask
            x[i, :n], t[i, :n], mask[i, :n] = b.flat[s : s + n], b.extras["flat_target"][s : s + n], True
        if args.packed:
            batches.append((x[mask].to(dev), t[mask].to(dev), None, b.total_rows))
        else:
            batches.append((x.to(dev), t.to(dev), mask.to(dev), b.total_rows))

    def step(x, t, mask, _):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = model(x)
        if mask is None:
            loss = torch.nn.functional.mse_loss(y.float(), t.float())
        else:
            loss = torch.nn.functional.mse_loss(y[mask].float(), t[mask].float())
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for i in range(max(args.warmup, 3)):
        step(*batches[i % NUM_BATCHES])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tokens = 0
    e0.record()
    for i in range(args.steps):
        step(*batches[i % NUM_BATCHES])
        tokens += batches[i % NUM_BATCHES][3]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"impl": "eager_pytorch_bar", "layout": "packed" if args.packed else "padded [B, L_max, Din]",
                      "metric": "aligner_train_tokens_per_sec", "value": tokens / (ms * 1e-3), "unit": "tokens/s", "n_gpus": 1,
                      "steps": args.steps, "ms_per_step": ms / args.steps, "dtype": "bf16 autocast",
                      "rows_computed_per_step": float(sum(b[0].shape[0] * (b[0].shape[1] if b[0].dim() == 3 else 1) for b in batches)) / NUM_BATCHES,
                      "valid_tokens_per_step": tokens / args.steps}))


if __name__ == "__main__":
    main()
