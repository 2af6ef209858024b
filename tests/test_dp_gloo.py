"""World-size-2 data-parallel logic on CPU (gloo): sharding + flat gradient buckets + pre-scaled SUM all-reduce must
equal DDP's semantics -- the mean over ranks of each rank's own mean-loss gradients (runner_base.py:88-92) -- computed
here by the oracle in one process. The CUDA kernels are not involved (no GPU here); gradients come from the oracle."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist

from _mp import spawn_ranks
from oracle import aligner_ref
from thinkdiff_mlre_b200.aligner import DataParallelState, GradBuckets
from thinkdiff_mlre_b200.sharding import shard_bounds

DIN, D, SEQS = 64, 128, 5
NAMES = ("dW1", "db1", "dW2", "db2", "dg")


def _global_batch():
    rng = np.random.RandomState(42)
    lens = [3, 9, 4, 7, 2]  # ragged: the two shards hold different token counts
    xs = [torch.from_numpy(rng.standard_normal((n, DIN)).astype(np.float32)) for n in lens]
    ts = [torch.from_numpy(rng.standard_normal((n, D)).astype(np.float32)) for n in lens]
    return xs, ts


def _shard_grads(xs, ts, params):
    x, t = torch.cat(xs), torch.cat(ts)
    fwd = aligner_ref.aligner_fwd_bwd_manual(x, params, regime="fp32")
    dy = 2.0 * (fwd["y"] - t) / t.numel()  # this rank's OWN mean loss
    out = aligner_ref.aligner_fwd_bwd_manual(x, params, dy=dy, regime="fp32")
    return [out[n] for n in NAMES]


def _worker(rank, world, init, overlap, ret):
    dist.init_process_group("gloo", init_method=init, rank=rank, world_size=world)
    try:
        params = aligner_ref.init_params_numpy(DIN, D, seed=3)
        xs, ts = _global_batch()
        lo, hi = shard_bounds(SEQS, world, rank)
        grads = _shard_grads(xs[lo:hi], ts[lo:hi], params)
        dp = DataParallelState(None, overlap=overlap)
        gb = GradBuckets(DIN, D, "cpu")
        for dst, src in zip(gb.in_parameter_order(), grads):
            dst.copy_(src / dp.world)  # the kernels write gradients pre-scaled by 1/world
        works = [dp.all_reduce_async(gb.linear2), dp.all_reduce_async(gb.linear1)]
        for w in works:
            w.wait()
        if rank == 0:
            ret.put([g.clone().numpy() for g in gb.in_parameter_order()])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_gradient_mean_matches_oracle(overlap):
    world = 2
    got = spawn_ranks(_worker, world, (overlap,), results=1)[0]
    params = aligner_ref.init_params_numpy(DIN, D, seed=3)
    xs, ts = _global_batch()
    per_rank = [_shard_grads(xs[slice(*shard_bounds(SEQS, world, r))], ts[slice(*shard_bounds(SEQS, world, r))], params) for r in range(world)]
    for i, name in enumerate(NAMES):
        want = sum(g[i] for g in per_rank) / world
        np.testing.assert_allclose(got[i], want.numpy(), rtol=1e-5, atol=1e-8, err_msg=name)
    # DDP's mean-of-means is NOT the global token mean when shards are ragged -- make sure we kept DDP's
    glob = _shard_grads(xs, ts, params)
    assert np.abs(got[2] - glob[2].numpy()).max() > 1e-6


def test_bucket_views_alias_flat_storage():
    gb = GradBuckets(DIN, D, "cpu")
    gb.linear2.zero_(), gb.linear1.zero_()
    gb.dW2.fill_(1), gb.db2.fill_(2), gb.dg.fill_(3), gb.dW1.fill_(4), gb.db1.fill_(5)
    assert gb.linear2.numel() == D * D + 2 * D and gb.linear1.numel() == D * DIN + D
    assert float(gb.linear2.sum()) == D * D + 2 * D + 3 * D and float(gb.linear1.sum()) == 4 * D * DIN + 5 * D
    assert [t.shape for t in gb.in_parameter_order()] == [(D, DIN), (D,), (D, D), (D,), (D,)]


def _peer_setup_worker(rank, world, init, ret):
    dist.init_process_group("gloo", init_method=init, rank=rank, world_size=world)
    try:
        from thinkdiff_mlre_b200.peer import PeerExchange, PeerSetupError

        try:
            PeerExchange(192, 512, None, torch.device("cuda", 0))
            ret.put((rank, "no error"))
        except PeerSetupError as e:
            ret.put((rank, str(e)))
    finally:
        dist.destroy_process_group()


def test_peer_setup_fails_on_every_rank_with_one_message():
    """No GPU here, so allocating the exchange buffer fails on both ranks: the set-up is collective and must raise the SAME
    PeerSetupError everywhere (bench.py's `--dp auto` then falls back to the NCCL exchange on all ranks) instead of leaving a rank
    blocked in the handle exchange."""
    if torch.cuda.is_available():
        pytest.skip("exercises the failure path of a box without CUDA")
    got = dict(spawn_ranks(_peer_setup_worker, 2))
    assert got[0] == got[1] and got[0].startswith("peer data parallel set-up failed") and "rank 0" in got[0] and "rank 1" in got[0]


def _dying_worker(rank, world, init, ret):
    dist.init_process_group("gloo", init_method=init, rank=rank, world_size=world)
    if rank == 0:
        raise SystemExit(3)  # dies before reporting; rank 1 would wait for it forever
    dist.barrier()
    ret.put(rank)


def test_spawn_ranks_reports_a_dead_rank_at_once_and_reaps_the_others():
    """The process harness itself: a rank that dies must fail the test within seconds (not after the result timeout) and must
    not leave its peer running -- a stuck orphan once kept a GPU box busy for 15 minutes after its test had already failed."""
    import time

    t0 = time.monotonic()
    with pytest.raises(RuntimeError, match="died before reporting"):
        spawn_ranks(_dying_worker, 2, timeout=120)
    assert time.monotonic() - t0 < 60


def test_bench_resolves_the_nvml_handle_by_uuid_not_by_index():
    """bench.py samples clocks / binds NUMA through NVML, which enumerates every GPU of a box while CUDA enumerates
    CUDA_VISIBLE_DEVICES: the handle must come from the CUDA device's UUID."""
    import importlib.util
    import types

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    calls = []
    fake = types.SimpleNamespace(nvmlDeviceGetHandleByUUID=lambda u: calls.append(("uuid", u)) or "H-uuid",
                                 nvmlDeviceGetHandleByPciBusId=lambda b: calls.append(("pci", b)) or "H-pci",
                                 nvmlDeviceGetHandleByIndex=lambda i: calls.append(("index", i)) or "H-index")
    real = torch.cuda.get_device_properties
    try:
        torch.cuda.get_device_properties = lambda i: types.SimpleNamespace(uuid="abc-123", pci_bus_id=7, pci_domain_id=0, pci_device_id=0)
        assert bench.nvml_handle(fake, 0) == "H-uuid" and calls == [("uuid", b"GPU-abc-123")]
        torch.cuda.get_device_properties = lambda i: (_ for _ in ()).throw(RuntimeError("no CUDA"))
        assert bench.nvml_handle(fake, 3) == "H-index"
    finally:
        torch.cuda.get_device_properties = real


def _sharded_checkpoint_worker(rank, world, init, ret):
    dist.init_process_group("gloo", init_method=init, rank=rank, world_size=world)
    try:
        import thinkdiff_mlre_b200 as td

        torch.manual_seed(0)
        m = td.ThinkDiffAligner(DIN, D)
        m.enable_data_parallel(defer_wait=True, sharded=True)
        opt = td.FusedAdamW(m, lr=1e-3)
        g = torch.Generator().manual_seed(5)  # the GLOBAL moments, identical on every rank; each rank keeps only its rows
        full = {}
        for name, p in m.named_parameters():
            full[name] = (torch.randn(p.shape, generator=g), torch.rand(p.shape, generator=g))
            ea, es = full[name]
            if p.dim() == 2:
                lo, hi = m._dp.shard_rows(p.shape[0])
                opt.state[p] = {"exp_avg": ea[lo:hi].clone(), "exp_avg_sq": es[lo:hi].clone(), "shard_rows": (lo, hi), "step": 9}
            else:
                opt.state[p] = {"exp_avg": ea.clone(), "exp_avg_sq": es.clone(), "step": 9}
        opt._t = 9
        sd = opt.state_dict()  # collective: the row blocks are all-gathered
        names = [n for n, _ in m.named_parameters()]
        index = {id(p): i for i, p in enumerate(q for grp in opt.param_groups for q in grp["params"])}
        gathered_ok = all(torch.equal(sd["state"][index[id(p)]]["exp_avg"], full[n][0]) and torch.equal(sd["state"][index[id(p)]]["exp_avg_sq"], full[n][1])
                          and "shard_rows" not in sd["state"][index[id(p)]] for n, p in m.named_parameters())
        still_sharded = all(opt.state[p]["exp_avg"].shape[0] == p.shape[0] // world for p in m.parameters() if p.dim() == 2)
        # torch.optim.AdamW accepts it (what a single-GPU resume of the checkpoint would do)
        plain = dict(sd)
        plain.pop("fused_adamw_step")
        from thinkdiff_mlre_b200.train_step import reference_param_groups

        torch.optim.AdamW(reference_param_groups(m, 0.05), lr=1e-3).load_state_dict(plain)
        # rank 0's checkpoint restored on EVERY rank (runner_base.py:662): each rank must end up with ITS rows, not rank 0's
        box = [sd if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        m2 = td.ThinkDiffAligner(DIN, D)
        m2.enable_data_parallel(defer_wait=True, sharded=True)
        opt2 = td.FusedAdamW(m2, lr=1e-3)
        opt2.load_state_dict(box[0])
        resliced_ok = opt2._t == 9
        for (n, p) in m2.named_parameters():
            st = opt2.state[p]
            if p.dim() == 2:
                lo, hi = m2._dp.shard_rows(p.shape[0])
                resliced_ok = resliced_ok and st["shard_rows"] == (lo, hi) and torch.equal(st["exp_avg"], full[n][0][lo:hi]) and torch.equal(st["exp_avg_sq"], full[n][1][lo:hi])
            else:
                resliced_ok = resliced_ok and torch.equal(st["exp_avg"], full[n][0])
            resliced_ok = resliced_ok and st["step"] == 9
        ret.put((rank, gathered_ok, still_sharded, resliced_ok, names))
    finally:
        dist.destroy_process_group()


def test_sharded_optimizer_checkpoint_gathers_full_moments_and_reslices_per_rank():
    """in the sharded / peer modes a rank holds the AdamW moments of its own weight rows only. state_dict()
    must return full-shape moments on every rank (collective all-gather) in torch.optim.AdamW's format, and rank 0's checkpoint
    loaded on every rank (the reference's resume, runner_base.py:613 / :662) must leave each rank with the moments of ITS rows and
    the step counter of the checkpoint."""
    got = sorted(spawn_ranks(_sharded_checkpoint_worker, 2))
    assert [g[0] for g in got] == [0, 1]
    for rank, gathered_ok, still_sharded, resliced_ok, names in got:
        assert gathered_ok and still_sharded and resliced_ok, (rank, gathered_ok, still_sharded, resliced_ok)
        assert names == ["0.weight", "0.bias", "2.weight", "2.bias", "3.weight"]


def _row_gather_worker(rank, world, init, ret):
    dist.init_process_group("gloo", init_method=init, rank=rank, world_size=world)
    try:
        import thinkdiff_mlre_b200 as td

        torch.manual_seed(0)
        m = td.ThinkDiffAligner(DIN, D)
        m.enable_data_parallel(defer_wait=True, sharded=True)
        want = {k: v.clone() for k, v in m.state_dict().items()}
        # every rank is current on ITS rows only (what the row-sharded AdamW leaves behind); the other rows are stale
        for w in (m[0].weight.data, m[2].weight.data):
            lo, hi = m._dp.shard_rows(w.shape[0])
            stale = torch.full_like(w, float("nan"))
            stale[lo:hi] = w[lo:hi]
            w.copy_(stale)
        m.sync_parameters()  # collective, in place
        ok = all(torch.equal(v, want[k]) for k, v in m.state_dict().items())
        ret.put((rank, ok, m._dp.shard_rows(D)))
    finally:
        dist.destroy_process_group()


def test_sync_parameters_all_gathers_the_master_rows_in_place():
    """Sharded data parallel: after the row-sharded updates each rank holds current fp32 master rows for its own block only;
    sync_parameters() (called by AlignerTrainStep.flush() before a checkpoint) must rebuild the full matrices on every rank."""
    got = sorted(spawn_ranks(_row_gather_worker, 2))
    assert [(g[0], g[1]) for g in got] == [(0, True), (1, True)]
    assert [g[2] for g in got] == [(0, D // 2), (D // 2, D)]
