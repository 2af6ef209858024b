"""GPU test of the shard-fed path: a batch read from flat on-disk shards by EmbedShardReader, shipped with
the overlapped prefetch and trained on, gives the same bits as the same batch handed over from memory; the reader's pinned ring
slots are only recycled after the copies that read them."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

DIN, D = 192, 512


def test_shard_fed_steps_equal_memory_fed_steps(tmp_path):
    import thinkdiff_mlre_b200 as td
    from oracle import aligner_ref

    m1 = td.ThinkDiffAligner(DIN, D).cuda()
    m1.load_state_dict(aligner_ref.init_params_numpy(DIN, D, seed=4))
    m2 = copy.deepcopy(m1)
    batches = [td.synthetic_lvlm_batch(6, 70, DIN, D, seed=300 + j, pin=False, truncated=True) for j in range(5)]
    fpath, tpath = str(tmp_path / "feat.tdemb"), str(tmp_path / "tgt.tdemb")
    with td.EmbedShardWriter(fpath, DIN) as wf, td.EmbedShardWriter(tpath, D) as wt:
        for b in batches:
            for s0, n in zip(b.src_row_start.tolist(), b.lens.tolist()):
                wf.add(b.flat[s0 : s0 + n], list(range(n)))
                wt.add(b.extras["flat_target"][s0 : s0 + n], list(range(n)))
    rf, rt = td.EmbedShardReader(fpath), td.EmbedShardReader(tpath)
    bi = dict(use_output_embed=1, use_input_embed=0, random_split_output_embed=0, output_embed_max_len=1 << 30, input_embed_max_len=1 << 30)
    s1 = td.AlignerTrainStep(m1, td.FusedAdamW(m1, lr=1e-3))
    s2 = td.AlignerTrainStep(m2, td.FusedAdamW(m2, lr=1e-3))
    losses1, losses2, handles = [], [], []
    for j, b in enumerate(batches):
        losses1.append(s1.step_host(b, "cuda"))
        fb = rf.batch(6 * j, 6 * (j + 1), bi)     # 5 batches through a 3-slot ring: slots are recycled
        tb = rt.batch(6 * j, 6 * (j + 1), bi)
        assert fb.flat.is_pinned() and "_h2d_enqueued" in fb.extras
        cbs = [fb.extras["_h2d_enqueued"], tb.extras["_h2d_enqueued"]]
        fb.extras["flat_target"] = tb.flat
        fb.extras["_h2d_enqueued"] = lambda evs, cbs=cbs: [cb(evs) for cb in cbs]
        assert fb.lens.tolist() == b.lens.tolist() and torch.equal(fb.flat, b.flat)
        handles.append(s2.prefetch(fb, "cuda"))   # never synchronised by the host: only the ring's events protect the slots
        losses2.append(s2.step_prefetched(handles[-1]))
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(losses1, losses2))
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.equal(a, b)
    rf.close(), rt.close()
