"""Peer-memory data parallel (thinkdiff_mlre_b200/peer.py): the gradient exchange fused into the weight-gradient GEMM epilogues.

The kernels take plain pointer arrays, so everything except the CUDA-IPC mapping is exercised on ONE GPU: the "peers" of the
loop-back tests are separate buffers of the same device. The two-process test needs 2 GPUs (gpurun --gpus 2); bench.py repeats
its check (`dp_parity`) at every N > 1."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DIN, D = 192, 512


def _arr(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


@pytest.mark.parametrize("stream_k", [False, True])
@pytest.mark.parametrize("world", [1, 2, 8])
@pytest.mark.parametrize("shape", [(512, 192, 300), (512, 512, 1000), (4096, 4096, 2048)])
def test_scatter_gemm_matches_local_gemm_bit_exactly(world, shape, stream_k):
    from thinkdiff_mlre_b200 import _lib as L
    from thinkdiff_mlre_b200 import ops

    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((K, M), generator=g, device="cuda").to(torch.bfloat16)  # MN-major operands: [K, rows]
    B = torch.randn((K, N), generator=g, device="cuda").to(torch.bfloat16)
    want = ops.gemm_f32out(A, B, True, True, alpha=0.5, cta_pair=True, stream_k=stream_k)
    parts = [torch.full((M // world, N), float("nan"), device="cuda") for _ in range(world)]
    ws, ws_bytes = ops.gemm_workspace(A.device, stream_k)
    L.check(L.lib().td_gemm_tn_scatter(L.ptr(A), M, L.ptr(B), N, M, N, K, 0.5, _arr(parts), world, L.ptr(ws), ws_bytes, L.stream_ptr()),
            "td_gemm_tn_scatter")
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(parts), want)  # same schedule, same accumulation order: same bits as the local-output GEMM


def test_flags_post_sum_and_adamw_slots_loopback():
    from thinkdiff_mlre_b200 import _lib as L

    world, n = 4, 4096
    flags = [torch.zeros(64, dtype=torch.int32, device="cuda") for _ in range(world)]
    # signal: +1 on counter `slot` of every flag array (one call per "rank" and step); wait: the local counter has reached
    # steps * world, i.e. every rank has signalled every step
    steps = 3
    for _ in range(steps):
        for _src in range(world):
            L.check(L.lib().td_peer_signal(_arr(flags), world, 32, L.stream_ptr()), "td_peer_signal")
    L.check(L.lib().td_peer_wait(C.c_void_p(flags[2].data_ptr() + 4 * 32), 1, steps * world, 5.0, L.stream_ptr()), "td_peer_wait")
    torch.cuda.synchronize()
    for f in flags:
        assert int(f[32]) == steps * world and int(f.sum()) == steps * world
    # post + sum (fixed order)
    g = torch.Generator(device="cuda").manual_seed(1)
    srcs = [torch.randn(n, generator=g, device="cuda") for _ in range(world)]
    slots = torch.zeros((world, n), device="cuda")
    for s, src in enumerate(srcs):
        L.check(L.lib().td_peer_post(L.ptr(src), _arr([slots[s]]), 1, n, L.stream_ptr()), "td_peer_post")
    out = torch.empty(n, device="cuda")
    L.check(L.lib().td_sum_slots(L.ptr(slots), n, world, L.ptr(out), n, L.stream_ptr()), "td_sum_slots")
    want = srcs[0].clone()
    for s in srcs[1:]:
        want += s
    assert torch.equal(out, want)
    # AdamW over slots == td_adamw_step on the pre-summed gradient; bf16 rows land in every destination
    p0 = torch.randn(n, generator=g, device="cuda")
    pa, pb = p0.clone(), p0.clone()
    ma, va, mb, vb = (torch.zeros(n, device="cuda") for _ in range(4))
    dst = [torch.zeros(n, dtype=torch.bfloat16, device="cuda") for _ in range(3)]
    ref_bf16 = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    for step in (1, 2):
        L.check(L.lib().td_adamw_slots_step(L.ptr(pa), L.ptr(slots), n, world, L.ptr(ma), L.ptr(va), _arr(dst), 3, n, 0.05, 1e-3, 0.9,
                                            0.999, 1e-8, step, 0.5, None, L.stream_ptr()), "td_adamw_slots_step")
        one = lambda t: (C.c_void_p * 1)(t.data_ptr())  # noqa: E731
        L.check(L.lib().td_adamw_step(1, one(pb), one(want), one(mb), one(vb), one(ref_bf16), (C.c_int64 * 1)(n), (C.c_float * 1)(0.05),
                                      1e-3, 0.9, 0.999, 1e-8, step, 0.5, None, L.stream_ptr()), "td_adamw_step")
    torch.cuda.synchronize()
    assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb)
    for t in dst:
        assert torch.equal(t, ref_bf16)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _train(rank, world, port, mode, ret):
    import torch.distributed as dist

    import thinkdiff_mlre_b200 as td
    from oracle import aligner_ref

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        m = td.ThinkDiffAligner(DIN, D).cuda()
        m.load_state_dict(aligner_ref.init_params_numpy(DIN, D, seed=3))
        m.enable_data_parallel(defer_wait=True, sharded=mode != "plain", peer=mode == "peer")
        step = td.AlignerTrainStep(m, td.FusedAdamW(m, lr=1e-3), pipelined=True)
        losses = []
        for j in range(4):
            b = td.synthetic_lvlm_batch(4, 50, DIN, D, seed=10 * j + rank, pin=False)
            losses.append(step.step_device(b.flat.cuda(), b.src_row_start.cuda(), b.lens.cuda(), b.total_rows, b.l_max,
                                           b.extras["flat_target"].cuda()))
        step.flush()
        torch.cuda.synchronize()
        ret.put((rank, [p.detach().float().cpu().numpy() for p in m.parameters()], [float(l) for l in losses]))
    finally:
        dist.destroy_process_group()


def _run(world, mode):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train, args=(r, world, port, mode, ret)) for r in range(world)]
    for p in procs:
        p.start()
    res = [ret.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return {r: (params, losses) for r, params, losses in res}


def test_peer_training_world1_equals_plain_pipeline():
    """Degenerate single-rank case: every kernel of the peer path runs (scatter epilogue, flags, slot AdamW), no IPC."""
    plain, peer = _run(1, "plain"), _run(1, "peer")
    for a, b in zip(plain[0][0], peer[0][0]):
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(plain[0][1], peer[0][1], rtol=1e-6)


def test_peer_training_two_gpus_equals_sharded_nccl_and_keeps_replicas_identical():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    sharded, peer = _run(2, "sharded"), _run(2, "peer")
    for a, b in zip(peer[0][0], peer[1][0]):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(sharded[0][0], peer[0][0]):
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(sharded[0][1], peer[0][1], rtol=1e-6)
