"""Process harness of the multi-rank tests (gloo on CPU, NCCL / peer memory on GPUs).

* Rendezvous through a FILE store (``init_method="file://..."``): no TCP port to pick, so no race with whatever else runs on the
  box (a "free" port found by bind-and-release can be taken again before rank 0 listens on it -- that killed rank 0 once and left
  rank 1 waiting for it forever).
* A rank that dies is noticed at once (its exit code), and EVERY child is killed before the helper returns, pass or fail: a
  crashed rank can never leave its peers -- or pytest's own exit -- hanging.
"""
import os
import queue
import tempfile
import time

import torch.multiprocessing as mp


def spawn_ranks(target, world, args=(), results=None, timeout=150.0):
    """Run ``target(rank, world, init_method, *args, ret)`` in ``world`` spawned processes; ``ret`` is a queue the ranks put their
    results on. Returns ``results`` (default ``world``) items in arrival order."""
    results = world if results is None else results
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    fd, path = tempfile.mkstemp(prefix="td_rdzv_")
    os.close(fd)
    os.remove(path)  # the FileStore creates it
    procs = [ctx.Process(target=target, args=(r, world, f"file://{path}", *args, ret), daemon=True) for r in range(world)]
    got = []
    try:
        for p in procs:
            p.start()
        deadline = time.monotonic() + timeout
        while len(got) < results:
            try:
                got.append(ret.get(timeout=0.5))
                continue
            except queue.Empty:
                pass
            dead = [(i, p.exitcode) for i, p in enumerate(procs) if p.exitcode not in (None, 0)]
            if dead:
                raise RuntimeError(f"rank(s) died before reporting: {dead} (see captured stderr)")
            if time.monotonic() > deadline:
                raise TimeoutError(f"{results - len(got)} result(s) missing after {timeout:.0f} s")
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0, f"rank exit code {p.exitcode}"
        return got
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
        for p in procs:
            p.join(timeout=10)
        try:
            os.remove(path)
        except OSError:
            pass
