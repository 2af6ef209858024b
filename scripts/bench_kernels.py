#!/usr/bin/env python
"""Per-kernel micro-benchmark of the HBM-bound kernels (pack, norm, losses) at BASELINE sizes: achieved GB/s of
ALGORITHMIC bytes vs the measured HBM copy peak. CUDA events, L2 flushed between iterations, 20 iterations after 3 warm-ups.
    python scripts/bench_kernels.py [--out profiles/r01_kernel_microbench.json]
Run the same command under `ncu --set full` for the per-kernel captures committed in profiles/."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import thinkdiff_mlre_b200 as td  # noqa: E402
from thinkdiff_mlre_b200 import ops  # noqa: E402


def timed(fn, flush, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        flush.zero_()  # 256 MB write: evicts the 126 MB L2
        e0[i].record()
        fn()
        e1[i].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in zip(e0, e1))
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}

    def report(name, ms, nbytes, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GB/s": gbs, "frac_of_measured_hbm": gbs / peaks["hbm_gbs"], "what": note}
        print(f"{name:28s} {ms*1e3:8.1f} us  {gbs:7.0f} GB/s  {gbs/peaks['hbm_gbs']:.2f} of measured HBM   {note}")

    # pack: config 5 per-GPU shard (128 sequences, len <= 1024, d = 3584) and config 4 composition rows (d = 4096)
    for name, B, L, C in (("pack_cfg5_3584", 128, 1024, 3584), ("pack_cfg2_3584", 64, 256, 3584), ("pack_cfg4_4096", 256, 128, 4096)):
        b = td.synthetic_lvlm_batch(B, L, C, 64, seed=1, pin=False, with_target=False)
        flat, start, lens = b.flat.to(dev), b.src_row_start.to(dev), b.lens.to(dev)
        cu = ops.cu_seqlens(lens)
        ms = timed(lambda: ops.pack_varlen(flat, start, cu, b.total_rows), flush)
        report(name, ms, 2 * b.total_rows * C * 2, f"{b.total_rows} rows x {C} bf16, read + write")
        if name == "pack_cfg5_3584":
            ms = timed(lambda: ops.pack_padded(flat, start, cu, b.l_max), flush)
            report("pack_padded_cfg5", ms, (b.total_rows + B * b.l_max) * C * 2 + B * b.l_max * 8, "reference layout: zero pad + int64 mask")
    M, D = 65536, 4096  # config 5 token count per GPU
    h2 = torch.randn(M, D, device=dev).to(torch.bfloat16)
    g = torch.ones(D, device=dev)
    ms = timed(lambda: ops.rmsnorm_fwd(h2, g), flush)
    report("rmsnorm_fwd_fp32out", ms, M * D * 6, "bf16 in, fp32 out")
    y, rstd = ops.rmsnorm_fwd(h2, g)
    dy = torch.randn(M, D, device=dev)
    ms = timed(lambda: ops.rmsnorm_bwd(dy, h2, rstd, g), flush)
    report("rmsnorm_bwd_fp32dy", ms, M * D * 8, "fp32 dy + bf16 x in, bf16 dx out")
    t = torch.randn(M, D, device=dev).to(torch.bfloat16)
    ms = timed(lambda: ops.masked_mse_fwd_bwd(y, t), flush)
    report("masked_mse_fp32y_bf16t", ms, M * D * 10, "fp32 y + bf16 t in, fp32 dy out")
    y16 = y.to(torch.bfloat16)
    ms = timed(lambda: ops.masked_mse_fwd_bwd(y16, t), flush)
    report("masked_mse_bf16", ms, M * D * 6, "bf16 y, t, dy (SURVEY figure: 24 576 B/token)")
    del y, dy, y16
    R, V = 8192, 32128  # 64 x 128 decoder positions, Flan-T5-XXL vocabulary
    z = torch.randn(R, V, device=dev).to(torch.bfloat16)
    labels = torch.randint(0, V, (R,), device=dev)
    labels[::5] = -100
    ms = timed(lambda: ops.masked_ce_fwd_bwd(z, labels), flush)
    report("masked_ce_bf16_V32128", ms, int(0.8 * R) * V * 4 + int(0.2 * R) * V * 2, "bf16 logits in, bf16 dlogits out (ignored rows: zeros written only)")
    if args.out:
        json.dump({"peaks": peaks, "kernels": res}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
