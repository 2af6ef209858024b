"""Developer aid for ncu: a few sequential (single-stream) cfg-2 train steps at N=1 -- fixed launch order per step:
cu_seqlens, pack, GEMM fwd1, GEMM fwd2, norm+MSE+norm-bwd, finisher, GEMM dW2, GEMM dh0, finisher, GEMM dW1, AdamW x2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import thinkdiff_mlre_b200 as td

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = td.ThinkDiffAligner(3584, 4096).to(dev)
opt = td.FusedAdamW(m, lr=1e-4)
step = td.AlignerTrainStep(m, opt, pipelined=False)
bs = [td.synthetic_lvlm_batch(64, 256, 3584, 4096, seed=1234 + 1000 * j, pin=False) for j in range(2)]
res = [(b.flat.to(dev), b.src_row_start.to(dev), b.lens.to(dev), b.total_rows, b.l_max, b.extras["flat_target"].to(dev)) for b in bs]
for i in range(steps):
    loss = step.step_device(*res[i % 2])
torch.cuda.synchronize()
print("loss", float(loss))
