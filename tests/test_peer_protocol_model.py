"""Model check of the flag protocol behind the peer-memory data-parallel mode (thinkdiff_mlre_b200/peer.py).

The device path cannot run here, but its ORDERING argument can: this test replays, under many random interleavings, the exact
sequence of operations every rank enqueues on its two streams (compute: ``AlignerTrainStep._step_pipelined_peer`` +
``aligner._peer_backward``; update: ``FusedAdamW.launch_peer_update`` / ``launch_peer_small_update``) against a small memory
model, and asserts the three hazards can never happen:

  * an owner sums a gradient slot that a peer has not finished writing for this step, or is already overwriting for the next;
  * a rank's GEMM reads weight rows while their owner is storing newer ones into them, or reads rows that are a step stale;
  * the small-vector slots are summed while being re-posted.

Every data operation is split into begin / end so that overlapping accesses are visible to the checker. Every flag row is ONE
monotone counter per destination rank: a signal is a remote atomic add of 1 (release) that becomes visible only after the
signalling stream's earlier operations have ENDED (in-order streams); a wait for step ``t`` blocks its stream until the counter
reaches ``t * world``, i.e. until every rank has signalled ``t`` times. Two sync points per step and direction: GRAD1 (the small
vectors ride along: they are posted before the signal) and GRAD2 towards the owners, W1 (covers the replicated bias / norm
vectors, updated before the signal) and W2 back.
"""
import random

import pytest

GRAD1, GRAD2, W1, W2 = range(4)


class Violation(AssertionError):
    pass


class Memory:
    def __init__(self, world):
        self.world = world
        # version = step of the last completed write; writers / readers = accesses in flight
        self.slot = {(w, owner, src): {"v": 0, "writers": 0, "readers": 0} for w in (1, 2) for owner in range(world) for src in range(world)}
        self.small = {(dst, src): {"v": 0, "writers": 0, "readers": 0} for dst in range(world) for src in range(world)}
        self.weight = {(w, holder, owner): {"v": 0, "writers": 0, "readers": 0} for w in (1, 2) for holder in range(world) for owner in range(world)}
        self.vecs = {r: {"v": 0, "writers": 0, "readers": 0} for r in range(world)}  # b1 / b2 / g compute copies of rank r
        self.flags = {(row, dst): 0 for row in range(4) for dst in range(world)}  # counters: += 1 per signal

    @staticmethod
    def begin_read(cell, want, what):
        if cell["writers"]:
            raise Violation(f"{what}: read while a write is in flight")
        if cell["v"] != want:
            raise Violation(f"{what}: read version {cell['v']}, expected {want}")
        cell["readers"] += 1

    @staticmethod
    def end_read(cell, want, what):
        cell["readers"] -= 1
        if cell["writers"] or cell["v"] != want:
            raise Violation(f"{what}: data changed under a reader")

    @staticmethod
    def begin_write(cell, what):
        if cell["writers"] or cell["readers"]:
            raise Violation(f"{what}: write while another access is in flight")
        cell["writers"] += 1

    @staticmethod
    def end_write(cell, version):
        cell["writers"] -= 1
        cell["v"] = version


def compute_stream(rank, world, steps):
    """Operations of the compute stream of ``rank``: ("wait", row, step) | ("signal", row) | data operations."""
    ops = []
    for t in range(1, steps + 1):
        if t > 1:
            ops.append(("wait", W1, t - 1))
        ops.append(("read_weight", 1, t - 1))        # GEMM1
        ops.append(("read_vecs", t - 1))             # b1 in GEMM1's epilogue
        if t > 1:
            ops.append(("wait", W2, t - 1))
        ops.append(("read_weight", 2, t - 1))        # GEMM2
        ops.append(("read_vecs", t - 1))             # b2 / g in GEMM2's epilogue and the fused norm kernel
        ops.append(("read_weight", 2, t - 1))        # backward: dh0 = dh2 . W2
        ops.append(("post_small", t))                # [db2 | dg | db1] to every rank, right after the finisher
        ops.append(("write_slots", 1, t))            # dW1 GEMM, scatter epilogue
        ops.append(("signal", GRAD1))
        ops.append(("write_slots", 2, t))            # dW2 GEMM
        ops.append(("signal", GRAD2))
    return ops


def update_stream(rank, world, steps, vecs_before_signal=True):
    ops = []
    for t in range(1, steps + 1):
        ops += [("wait", GRAD1, t), ("adamw", 1, t)]
        small = [("sum_small", t), ("update_vecs", t)]
        ops += (small + [("signal", W1)]) if vecs_before_signal else ([("signal", W1)] + small)
        ops += [("wait", GRAD2, t), ("adamw", 2, t), ("signal", W2)]
    return ops


class Stream:
    def __init__(self, rank, ops):
        self.rank, self.ops, self.pc, self.in_flight = rank, ops, 0, None

    def done(self):
        return self.pc >= len(self.ops) and self.in_flight is None


def runnable(st, mem):
    if st.in_flight is not None:
        return True
    if st.pc >= len(st.ops):
        return False
    op = st.ops[st.pc]
    if op[0] == "wait":
        return mem.flags[(op[1], st.rank)] >= op[2] * mem.world
    return True


def advance(st, mem):
    """Run one half-operation (begin or end) of the stream's current op."""
    r, world = st.rank, mem.world
    if st.in_flight is not None:
        st.in_flight()
        st.in_flight = None
        st.pc += 1
        return
    op = st.ops[st.pc]
    kind = op[0]
    if kind == "wait":
        st.pc += 1
    elif kind == "signal":
        for dst in range(world):
            mem.flags[(op[1], dst)] += 1
        st.pc += 1
    elif kind == "read_weight":
        w, want = op[1], op[2]
        cells = [mem.weight[(w, r, o)] for o in range(world)]
        for o, c in enumerate(cells):
            mem.begin_read(c, want, f"rank {r} reads W{w} rows of owner {o}")
        st.in_flight = lambda: [mem.end_read(c, want, f"rank {r} W{w}") for c in cells]
    elif kind == "read_vecs":
        c, want = mem.vecs[r], op[1]
        mem.begin_read(c, want, f"rank {r} reads its bias / norm copies")
        st.in_flight = lambda: mem.end_read(c, want, f"rank {r} vecs")
    elif kind == "write_slots":
        w, t = op[1], op[2]
        cells = [mem.slot[(w, o, r)] for o in range(world)]
        for o, c in enumerate(cells):
            mem.begin_write(c, f"rank {r} stores dW{w} slot at owner {o} (step {t})")
        st.in_flight = lambda: [mem.end_write(c, t) for c in cells]
    elif kind == "post_small":
        t = op[1]
        cells = [mem.small[(dst, r)] for dst in range(world)]
        for dst, c in enumerate(cells):
            mem.begin_write(c, f"rank {r} posts small vectors to rank {dst} (step {t})")
        st.in_flight = lambda: [mem.end_write(c, t) for c in cells]
    elif kind == "sum_small":
        t = op[1]
        cells = [mem.small[(r, src)] for src in range(world)]
        for src, c in enumerate(cells):
            mem.begin_read(c, t, f"rank {r} sums small slot of rank {src} (step {t})")
        st.in_flight = lambda: [mem.end_read(c, t, f"rank {r} small") for c in cells]
    elif kind == "update_vecs":
        t, c = op[1], mem.vecs[r]
        mem.begin_write(c, f"rank {r} updates its bias / norm copies (step {t})")
        st.in_flight = lambda: mem.end_write(c, t)
    elif kind == "adamw":
        w, t = op[1], op[2]
        slots = [mem.slot[(w, r, src)] for src in range(world)]
        dsts = [mem.weight[(w, holder, r)] for holder in range(world)]
        for src, c in enumerate(slots):
            mem.begin_read(c, t, f"owner {r} sums dW{w} slot of rank {src} (step {t})")
        for holder, c in enumerate(dsts):
            mem.begin_write(c, f"owner {r} stores W{w} rows into rank {holder} (step {t})")

        def end():
            for c in slots:
                mem.end_read(c, t, f"owner {r} dW{w} slots")
            for c in dsts:
                mem.end_write(c, t)

        st.in_flight = end
    else:
        raise ValueError(kind)


def simulate(world, steps, seed, compute=compute_stream, update=update_stream):
    rng = random.Random(seed)
    mem = Memory(world)
    streams = [Stream(r, compute(r, world, steps)) for r in range(world)] + [Stream(r, update(r, world, steps)) for r in range(world)]
    # a biased scheduler: sometimes let one rank race far ahead of the others
    while not all(s.done() for s in streams):
        ready = [s for s in streams if runnable(s, mem)]
        if not ready:
            raise Violation("deadlock: every stream is blocked")
        if rng.random() < 0.3:
            fav = rng.randrange(world)
            pref = [s for s in ready if s.rank == fav]
            ready = pref or ready
        advance(rng.choice(ready), mem)
    return mem


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_no_hazard_and_no_deadlock_under_random_interleavings(world):
    for seed in range(150):
        mem = simulate(world, steps=4, seed=seed)
        assert all(c["v"] == 4 for c in mem.weight.values())
        assert all(c["v"] == 4 for c in mem.vecs.values())


def test_the_checker_catches_a_missing_wait():
    """Drop the wait on the W2 counter before GEMM2: some interleaving must read stale or in-flight rows."""

    def broken(rank, world, steps):
        return [op for op in compute_stream(rank, world, steps) if op[:2] != ("wait", W2)]

    with pytest.raises(Violation):
        for seed in range(300):
            simulate(3, steps=4, seed=seed, compute=broken)


def test_the_checker_catches_a_wrong_update_order():
    """Bias / norm vectors updated AFTER the W1 signal: the next forward may read them a step stale, or a peer may re-post its
    small slot while this rank still sums it."""

    def late_small(rank, world, steps):
        return update_stream(rank, world, steps, vecs_before_signal=False)

    with pytest.raises(Violation):
        for seed in range(600):
            simulate(3, steps=5, seed=seed, update=late_small)
