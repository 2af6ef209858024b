"""Regenerate tests/golden/*.npz by running the REFERENCE'S OWN code (dev container only; needs /root/reference).

    python -m oracle.make_golden

Each fixture stores seeded inputs (numpy RandomState, so they do not depend on the torch RNG stream) and the outputs of
the reference functions exec'd from source by ``oracle.ref_loader``:
  aligner_small_{fp32,bf16}.npz   build_vision_projector('mlp2x_gelu_t5_norm') fwd + MSE loss + backward, all tensors
  aligner_mid_{fp32,bf16}.npz     multi-tile dims (Din 192, D 512, 300 rows): y, loss, small grads, sampled dW entries
  aligner_cfg1_fp32.npz           BASELINE config 1 (4 x 32 x 768 -> 4096, fp32): y rows, loss, sampled grad entries
  collater_{random_split,fixed_max,input_embed}.npz   the reference collater's three branches on ragged batches
  ce_small.npz                    CrossEntropyLoss(ignore_index=-100) expression of ...embed_decoder_2.py:243-246
  lm_head_ce_small.npz            lm_head (bias-free Linear, frozen) + that loss under CPU bf16 autocast, ...embed_decoder_2.py:239-246
  lr_schedule.npz                 the two LR scheduler classes of thinkdiff/common/optims.py at fixed (epoch, step) points
                                  (``python -m oracle.make_golden lr`` regenerates only this one)
The bf16 fixtures run the reference module under ``torch.autocast('cpu', dtype=bfloat16)`` -- the CPU analogue of the
training regime (thinkdiff/tasks/base_task.py:237).
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

from . import ref_loader
from .aligner_ref import init_params_numpy

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ref_module(din, d, seed):
    m = ref_loader.build_reference_projector(din, d, "mlp2x_gelu_t5_norm")
    params = init_params_numpy(din, d, seed)
    m.load_state_dict(params)
    return m, params


def _run(m, x, t, autocast):
    m.zero_grad(set_to_none=True)
    if autocast:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            y = m(x)
            loss = torch.nn.functional.mse_loss(y, t)
    else:
        y = m(x)
        loss = torch.nn.functional.mse_loss(y, t)
    loss.backward()
    return y.detach(), loss.detach(), {k: p.grad.detach().clone() for k, p in m.named_parameters()}


def _inputs(shape_x, d, seed, heavy_tail=False):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal(shape_x).astype(np.float32)
    if heavy_tail:  # a few outlier channels, as LLM hidden states have
        ch = rng.choice(shape_x[-1], size=min(8, shape_x[-1]), replace=False)
        x[..., ch] *= 50.0
    t = rng.standard_normal(shape_x[:-1] + (d,)).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(t)


def make_aligner(name, din, d, shape_x, seed, autocast, full, heavy_tail=False, sample=4096):
    m, params = _ref_module(din, d, seed)
    x, t = _inputs(shape_x, d, seed + 1, heavy_tail)
    y, loss, grads = _run(m, x, t, autocast)
    out = {
        "din": din, "d": d, "seed": seed, "autocast_bf16": int(autocast), "heavy_tail": int(heavy_tail),
        "x_shape": np.asarray(shape_x), "loss": loss.float().numpy(), "y_dtype": str(y.dtype),
    }
    if full:
        out.update({"x": x.numpy(), "t": t.numpy(), "y": y.float().numpy()})
        out.update({"p_" + k: v.numpy() for k, v in params.items()})
        out.update({"g_" + k: v.float().numpy() for k, v in grads.items()})
    else:
        rng = np.random.RandomState(seed + 2)
        yf = y.float().numpy().reshape(-1, d)
        rows = np.sort(rng.choice(yf.shape[0], size=min(16, yf.shape[0]), replace=False))
        out["y_rows"], out["y_sel"] = rows, yf[rows]
        out["y_abs_sum"] = np.float64(np.abs(yf.astype(np.float64)).sum())
        for k, v in grads.items():
            g = v.float().numpy()
            if g.ndim == 1:
                out["g_" + k] = g
            else:
                idx = rng.choice(g.size, size=sample, replace=False)
                out["gi_" + k], out["gs_" + k] = idx, g.reshape(-1)[idx]
                out["gn_" + k] = np.float64(np.linalg.norm(g.astype(np.float64)))
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(f"{name}: loss={float(loss):.6f} y dtype {y.dtype}")


def _ragged(ids_list):
    flat = np.concatenate([np.asarray(i, dtype=np.int64) for i in ids_list]) if ids_list else np.zeros(0, np.int64)
    off = np.zeros(len(ids_list) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(i) for i in ids_list])
    return flat, off


def make_collater(name, build_info, full_lens, C, seed):
    collater = ref_loader.load_collater()
    rng = np.random.RandomState(seed)
    samples, bits = [], []
    for i, L in enumerate(full_lens):
        # arbitrary 16-bit patterns reinterpreted as bf16: the collater must move bytes, not values
        w = rng.randint(0, 0x7F80, size=(L, C)).astype(np.uint16)  # finite positive bf16 patterns
        w |= (rng.randint(0, 2, size=(L, C)).astype(np.uint16) << 15)
        bits.append(w)
        emb = torch.from_numpy(w.view(np.int16).copy()).view(torch.bfloat16)
        ids = [int(v) for v in rng.randint(0, 32000, size=L)]
        samples.append({
            "json": {"generated_text": f"sample {i}", "output_token_ids": ids},
            "model.norm.input_embed.pth": emb.clone(),
            "model.norm.output_embed.pth": emb,
        })
    random.seed(seed)
    out = collater(build_info, samples)
    res = {"full_lens": np.asarray(full_lens), "C": C, "seed": seed,
           "src_bits": np.concatenate(bits, axis=0),
           "src_ids_flat": _ragged([s["json"]["output_token_ids"] for s in samples])[0]}
    for k, v in build_info.items():
        res["bi_" + k] = int(v)
    if build_info["use_output_embed"]:
        res["out_embed_bits"] = out["model.norm.output_embed"].view(torch.int16).numpy().view(np.uint16)
        res["out_mask"] = out["output_embed_mask"].numpy()
        assert out["output_embed_mask"].dtype == torch.int64
        res["ids_flat"], res["ids_off"] = _ragged(out["output_token_ids"])
    if build_info["use_input_embed"]:
        res["in_embed_bits"] = out["model.norm.input_embed"].view(torch.int16).numpy().view(np.uint16)
        res["in_mask"] = out["input_embed_mask"].numpy()
    np.savez_compressed(os.path.join(GOLDEN, name), **res)
    print(f"{name}: keys {sorted(out.keys())}")


def make_ce(name, R, V, seed):
    # the reference expression, thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:243-246
    from torch.nn import CrossEntropyLoss

    rng = np.random.RandomState(seed)
    logits = torch.from_numpy((rng.standard_normal((R, V)) * 3).astype(np.float32)).requires_grad_(True)
    labels = torch.from_numpy(rng.randint(0, V, size=R).astype(np.int64))
    labels[rng.rand(R) < 0.3] = -100
    labels[0] = -100
    loss_fct = CrossEntropyLoss(ignore_index=-100)
    loss = loss_fct(logits.view(-1, logits.size(-1)), labels.view(-1))
    loss.backward()
    np.savez_compressed(os.path.join(GOLDEN, name), logits=logits.detach().numpy(), labels=labels.numpy(),
                        loss=loss.detach().numpy(), dlogits=logits.grad.numpy())
    print(f"{name}: loss={float(loss):.6f}")


def make_lm_head_ce(name, R, K, V, seed):
    # the reference expressions, thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:239 and :243-246, under bf16 autocast
    # (tasks/base_task.py:237) with the head frozen (:715-717)
    from torch.nn import CrossEntropyLoss

    rng = np.random.RandomState(seed)
    lm_head = torch.nn.Linear(K, V, bias=False)
    with torch.no_grad():
        lm_head.weight.copy_(torch.from_numpy((rng.standard_normal((V, K)) * 0.5).astype(np.float32)))
    lm_head.weight.requires_grad_(False)
    seq = torch.from_numpy(rng.standard_normal((R, K)).astype(np.float32)).to(torch.bfloat16).requires_grad_(True)
    labels = torch.from_numpy(rng.randint(0, V, size=R).astype(np.int64))
    labels[rng.rand(R) < 0.3] = -100
    labels[1] = -100
    with torch.autocast("cpu", dtype=torch.bfloat16):
        lm_logits = lm_head(seq)
        loss_fct = CrossEntropyLoss(ignore_index=-100)
        loss = loss_fct(lm_logits.view(-1, lm_logits.size(-1)), labels.view(-1))
    loss.backward()
    assert lm_logits.dtype == torch.bfloat16 and loss.dtype == torch.float32 and seq.grad.dtype == torch.bfloat16
    np.savez_compressed(os.path.join(GOLDEN, name), seq=seq.detach().float().numpy(), weight=lm_head.weight.detach().numpy(),
                        labels=labels.numpy(), logits=lm_logits.detach().float().numpy(), loss=loss.detach().numpy(),
                        dseq=seq.grad.float().numpy())
    print(f"{name}: loss={float(loss):.6f}")


LR_CASES = {
    # the schedule every shipped config uses (configs/*.yaml: lr_sched / init_lr / min_lr / warmup_lr / warmup_steps / ...)
    "cosine_shipped": ("linear_warmup_cosine_lr", dict(max_epoch=40, iters_per_epoch=5000, min_lr=8e-5, init_lr=1e-4,
                                                       warmup_start_lr=1e-6, warmup_steps=2000, decay_rate=None)),
    "cosine_warmup_spans_epochs": ("linear_warmup_cosine_lr", dict(max_epoch=3, iters_per_epoch=10, min_lr=1e-5, init_lr=1e-3,
                                                                   warmup_start_lr=1e-6, warmup_steps=25, decay_rate=None)),
    "cosine_no_warmup": ("linear_warmup_cosine_lr", dict(max_epoch=2, iters_per_epoch=7, min_lr=0.0, init_lr=5e-4)),
    "step_decay": ("linear_warmup_step_lr", dict(max_epoch=6, iters_per_epoch=9, min_lr=2e-5, init_lr=1e-4, decay_rate=0.5,
                                                  warmup_start_lr=1e-6, warmup_steps=5)),
}


def lr_points(kw):
    ipe, me = kw["iters_per_epoch"], kw["max_epoch"]
    steps = sorted({0, 1, 2, ipe // 3, ipe // 2, ipe - 2, ipe - 1} & set(range(ipe)))
    epochs = sorted({0, 1, 2, me // 2, me - 1} & set(range(me)))
    pts = [(e, s) for e in epochs for s in steps]
    ws = kw.get("warmup_steps", 0)
    for it in (ws - 1, ws, ws + 1):  # around the end of the warm-up
        if 0 <= it < ipe * me:
            pts.append((it // ipe, it % ipe))
    return pts


def make_lr_schedule(name="lr_schedule.npz"):
    """LR values produced by the reference's own scheduler classes (thinkdiff/common/optims.py) at fixed (epoch, step) points."""
    import types

    classes = ref_loader.load_lr_schedulers()
    out = {}
    for case, (sched, kw) in LR_CASES.items():
        opt = types.SimpleNamespace(param_groups=[{"lr": -1.0}, {"lr": -1.0}])
        obj = classes[sched](optimizer=opt, **kw)
        pts, lrs = lr_points(kw), []
        for e, s in pts:
            obj.step(cur_epoch=e, cur_step=s)
            assert opt.param_groups[0]["lr"] == opt.param_groups[1]["lr"]
            lrs.append(opt.param_groups[0]["lr"])
        out[case + ".points"] = np.asarray(pts, dtype=np.int64)
        out[case + ".lr"] = np.asarray(lrs, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(f"{name}: {sum(len(v) for k, v in out.items() if k.endswith('.lr'))} lr values")


def main():
    assert ref_loader.available(), "needs /root/reference"
    if len(sys.argv) > 1 and sys.argv[1] == "lr":  # only the (later-added) schedule fixture; the others stay as committed
        os.makedirs(GOLDEN, exist_ok=True)
        return make_lr_schedule()
    if len(sys.argv) > 1 and sys.argv[1] == "lm_head":  # only the (later-added) output-head fixture
        os.makedirs(GOLDEN, exist_ok=True)
        return make_lm_head_ce("lm_head_ce_small.npz", 40, 128, 992, 11)
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    make_aligner("aligner_small_fp32.npz", 64, 128, (3, 7, 64), 11, autocast=False, full=True)
    make_aligner("aligner_small_bf16.npz", 64, 128, (3, 7, 64), 11, autocast=True, full=True)
    make_aligner("aligner_mid_fp32.npz", 192, 512, (300, 192), 23, autocast=False, full=False)
    make_aligner("aligner_mid_bf16.npz", 192, 512, (300, 192), 23, autocast=True, full=False)
    make_aligner("aligner_mid_bf16_heavy.npz", 192, 512, (300, 192), 29, autocast=True, full=False, heavy_tail=True)
    make_aligner("aligner_cfg1_fp32.npz", 768, 4096, (4, 32, 768), 31, autocast=False, full=False)
    make_aligner("aligner_cfg1_bf16.npz", 768, 4096, (4, 32, 768), 31, autocast=True, full=False)
    lens = [9, 2, 17, 5, 33, 12]
    make_collater("collater_random_split.npz",
                  dict(use_input_embed=0, use_output_embed=1, random_split_output_embed=1, output_embed_max_split_len=16,
                       output_embed_max_len=64, input_embed_max_len=64), lens, 16, 101)
    make_collater("collater_fixed_max.npz",
                  dict(use_input_embed=0, use_output_embed=1, random_split_output_embed=0, output_embed_max_split_len=16,
                       output_embed_max_len=12, input_embed_max_len=64), lens, 16, 102)
    make_collater("collater_fixed_max_uncapped.npz",
                  dict(use_input_embed=0, use_output_embed=1, random_split_output_embed=0, output_embed_max_split_len=16,
                       output_embed_max_len=64, input_embed_max_len=64), lens, 16, 103)
    make_collater("collater_input_embed.npz",
                  dict(use_input_embed=1, use_output_embed=1, random_split_output_embed=1, output_embed_max_split_len=128,
                       output_embed_max_len=64, input_embed_max_len=10), lens, 16, 104)
    make_ce("ce_small.npz", 24, 512, 7)
    make_lm_head_ce("lm_head_ce_small.npz", 40, 128, 992, 11)
    make_lr_schedule()


if __name__ == "__main__":
    main()
