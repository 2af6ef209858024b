import cProfile, pstats, sys, os, io
sys.path.insert(0, "/root/repo")
import torch
import thinkdiff_mlre_b200 as td
dev = torch.device("cuda")
torch.manual_seed(0)
a = td.ThinkDiffAligner(3584, 4096).to(dev)
opt = td.FusedAdamW(a)
st = td.AlignerTrainStep(a, opt, pipelined=True)
host = [td.synthetic_lvlm_batch(64, 256, 3584, 4096, seed=1234 + j) for j in range(2)]
res = [(b.flat.to(dev), b.src_row_start.to(dev), b.lens.to(dev), b.total_rows, b.l_max, b.extras["flat_target"].to(dev)) for b in host]
for i in range(5): st.step_device(*res[i % 2])
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(200): st.step_device(*res[i % 2])
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); ps = pstats.Stats(pr, stream=s).sort_stats("tottime"); ps.print_stats(28); print(s.getvalue()[:6000])
