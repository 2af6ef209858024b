"""Developer diagnostic: step-by-step comparison of the device GradScaler loop with torch.amp.GradScaler + torch.optim.AdamW."""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import thinkdiff_mlre_b200 as td
from oracle import aligner_ref
from thinkdiff_mlre_b200.train_step import reference_param_groups

DIN, D = 192, 512
m1 = td.ThinkDiffAligner(DIN, D).cuda(); m1.load_state_dict(aligner_ref.init_params_numpy(DIN, D, seed=9)); m2 = copy.deepcopy(m1)
batches = []
for j in range(4):
    b = td.synthetic_lvlm_batch(5, 60, DIN, D, seed=50 + j, pin=False)
    batches.append((b.flat.cuda(), b.src_row_start.cuda(), b.lens.cuda(), b.total_rows, b.l_max, b.extras["flat_target"].cuda()))
use_scaler = len(sys.argv) < 2 or sys.argv[1] != "noscaler"
ref_scaler = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=2) if use_scaler else None
opt1 = torch.optim.AdamW(reference_param_groups(m1, 0.05), lr=1e-3, weight_decay=0.05)
sc = td.DeviceGradScaler(torch.device("cuda"), init_scale=1024.0, growth_interval=2, enabled=use_scaler)
opt2 = td.FusedAdamW(m2, lr=1e-3)
step = td.AlignerTrainStep(m2, opt2, grad_scaler=sc)
for i, batch in enumerate(batches):
    flat, start, lens, total, _, tgt = batch
    cu = td.ops.cu_seqlens(lens); x, idx = td.ops.pack_varlen(flat, start, cu, total, want_index=True)
    loss = m1.mse_loss_packed(x, tgt, idx)
    (ref_scaler.scale(loss) if use_scaler else loss).backward()
    g1 = [p.grad.clone() for p in m1.parameters()]
    if use_scaler:
        ref_scaler.step(opt1); ref_scaler.update()
    else:
        opt1.step()
    opt1.zero_grad()
    # device path, by hand so the gradients can be inspected
    sc.begin_step()
    loss2 = m2.mse_loss_backward_packed(x, tgt, idx, upstream=sc.scale_tensor)
    g2 = [p.grad.clone() for p in m2.parameters()]
    flats = m2._grad_flats
    sc.accumulate_stats(flats["linear1"], flats["linear2"])
    sc.update(opt2, 0.0)
    opt2.step(ctl=sc.ctl)
    torch.cuda.synchronize()
    gd = [float((a - b).abs().max() / (a.abs().max() + 1e-30)) for a, b in zip(g1, g2)]
    pd = [float((a - b).abs().max() / (a.abs().max() + 1e-30)) for a, b in zip(m1.parameters(), m2.parameters())]
    print(f"step {i}: loss {float(loss):.6f} {float(loss2):.6f} grad-diff {['%.1e' % v for v in gd]} param-diff {['%.1e' % v for v in pd]}")
    print("   ctl", sc.state(), "stats", sc.stats.tolist(), "ref scale", ref_scaler.get_scale() if use_scaler else None)
