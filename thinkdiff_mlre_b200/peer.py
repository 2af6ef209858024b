"""Data parallel over NVLink peer memory: the exchange buffers and flags behind ``enable_data_parallel(peer=True)``.

Reference being replaced: DDP's bucketed NCCL all-reduce of the aligner gradients and the replicated ``optimizer.step()``
(thinkdiff/runners/runner_base.py:88-92, :98-127; thinkdiff/tasks/base_task.py:247-258). Here no collective kernel runs on the
step at all:

  * rank ``o`` owns rows ``[o D/N, (o+1) D/N)`` of W1 and W2 (as in the ZeRO-1 style ``sharded`` mode);
  * every rank's weight-gradient GEMM stores those rows of its gradient straight into ``slot[rank]`` of rank ``o``'s exchange
    buffer from its epilogue (``td_aligner_bwd_dh2_scatter``: 128-byte NVLink stores, un-split GEMM, no atomics);
  * rank ``o`` sums the N slots in rank order inside its AdamW pass (``td_adamw_slots_step``: bit-reproducible, and every
    replica is identical by construction) and stores the updated bf16 rows into every rank's compute copy of the weight;
  * the three small vectors are posted to every rank (``td_peer_post``) and summed there in the same order;
  * ordering is by counters in the destination's memory: ``td_peer_signal`` adds 1 to a row's counter at every rank (remote
    atomic, release), ``td_peer_wait`` lets a stream wait -- as a stream memory operation, no kernel -- until the counter shows
    that all ranks have signalled the step. Two sync points per step in each direction (GRAD1 / GRAD2 towards the owners,
    W1 / W2 back); the protocol is model-checked in tests/test_peer_protocol_model.py.

One exchange buffer per rank (``td_peer_alloc``: cudaMalloc + CUDA IPC handle; handles travel through
``torch.distributed.all_gather_object``, which is plumbing). ``ExchangeLayout`` is pure arithmetic and identical on all ranks.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

from . import _lib as L

MAX_PEERS = 8
FLAG_BYTES = 4096
# flag array (int32): one monotone counter per row, 128 bytes apart. Every rank adds 1 per step to the row's counter at every
# destination, so "all ranks have signalled step t" is the single comparison counter >= t * world.
#   GRAD1: dW1's rows AND the small vectors [db2 | dg | db1] of step t have been stored at their destinations
#   GRAD2: dW2's rows;  W1: the owner's rows of W1 (and, locally, b1 / b2 / g) are updated;  W2: the owner's rows of W2
ROW_GRAD1, ROW_GRAD2, ROW_W1, ROW_W2 = 0, 1, 2, 3
FLAG_ROW_STRIDE = 32  # int32 elements between rows (128 bytes: one line per row)


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


@dataclass(frozen=True)
class ExchangeLayout:
    """Byte offsets inside one rank's exchange buffer (the same on every rank)."""

    world: int
    din: int
    d: int

    def __post_init__(self):
        if not 1 <= self.world <= MAX_PEERS:
            raise ValueError(f"peer data parallel supports 1..{MAX_PEERS} ranks, got {self.world}")
        if self.d % self.world:
            raise ValueError(f"peer data parallel needs the {self.d} weight rows to divide by world size {self.world}")

    @property
    def rows(self) -> int:  # weight rows owned by one rank
        return self.d // self.world

    @property
    def small_numel(self) -> int:  # [db2 | dg | db1], padded to a multiple of 4 floats
        return _align(3 * self.d, 4)

    # -- gradient slots: slot s = the rows this rank owns, as computed by rank s; [world, rows, cols] fp32
    @property
    def off_g1(self) -> int:
        return FLAG_BYTES

    @property
    def slot1_numel(self) -> int:
        return self.rows * self.din

    @property
    def off_g2(self) -> int:
        return _align(self.off_g1 + 4 * self.world * self.slot1_numel)

    @property
    def slot2_numel(self) -> int:
        return self.rows * self.d

    @property
    def off_small(self) -> int:
        return _align(self.off_g2 + 4 * self.world * self.slot2_numel)

    # -- bf16 compute copies of the full weights (every owner writes its rows into every rank's copy)
    @property
    def off_w1(self) -> int:
        return _align(self.off_small + 4 * self.world * self.small_numel)

    @property
    def off_w2(self) -> int:
        return _align(self.off_w1 + 2 * self.d * self.din)

    @property
    def total_bytes(self) -> int:
        return _align(self.off_w2 + 2 * self.d * self.d)

    def flag_offset(self, row: int) -> int:
        return 4 * row * FLAG_ROW_STRIDE

    def grad_slot_offset(self, which: int, src: int) -> int:
        """Byte offset of slot ``src`` of weight ``which`` (1 or 2) inside the owner's buffer."""
        base, n = (self.off_g1, self.slot1_numel) if which == 1 else (self.off_g2, self.slot2_numel)
        return base + 4 * src * n

    def small_slot_offset(self, src: int) -> int:
        return self.off_small + 4 * src * self.small_numel

    def weight_rows_offset(self, which: int, owner: int) -> int:
        """Byte offset of owner's row block inside a rank's bf16 copy of weight ``which``."""
        base, cols = (self.off_w1, self.din) if which == 1 else (self.off_w2, self.d)
        return base + 2 * owner * self.rows * cols


class _DeviceBytes:
    """``__cuda_array_interface__`` view of raw device memory, so torch can alias it without a copy."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self._owner = owner  # keeps the allocation alive as long as any tensor aliases it


class PeerSetupError(RuntimeError):
    """Raised on EVERY rank when the exchange buffers cannot be allocated or mapped on some rank (no peer access, CUDA IPC not
    permitted in this container, out of memory): nothing is left mapped, the NCCL exchange modes remain usable."""


class PeerExchange:
    """This rank's exchange buffer + the mapped buffers of the other ranks, and the five tiny device operations on them."""

    def __init__(self, din: int, d: int, group=None, device=None, timeout_s: float = 600.0):
        import torch
        import torch.distributed as dist

        self.torch = torch
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.layout = ExchangeLayout(self.world, din, d)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.timeout_s = float(timeout_s)
        lay = self.layout
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        self._torch_backing = None
        self._local, self._opened = 0, []
        # Set-up is COLLECTIVE and must fail on all ranks or on none: a rank whose allocation / mapping fails still takes part in
        # the exchanges below, then every rank raises the same PeerSetupError (callers may fall back to the NCCL exchange).
        err = None
        try:
            if self.world == 1 and os.environ.get("TD_PEER_NO_IPC") == "1":
                # developer A/B: the same layout in ordinary caching-allocator memory (single rank only: nothing to export)
                self._torch_backing = torch.zeros((lay.total_bytes,), dtype=torch.uint8, device=self.device)
                ptr = C.c_void_p(self._torch_backing.data_ptr())
            else:
                with torch.cuda.device(self.device):
                    L.check(L.lib().td_peer_alloc(lay.total_bytes, C.byref(ptr), handle), "td_peer_alloc")
            self._local = int(ptr.value)
        except Exception as e:  # noqa: BLE001  (reported collectively below)
            err = f"rank {self.rank}: {e}"
        self.base = [0] * self.world  # base[o] = address of rank o's buffer in THIS process
        self.base[self.rank] = self._local
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, None if err else bytes(handle.raw), group=group)
            if err is None and all(h is not None for h in handles):
                try:
                    with torch.cuda.device(self.device):
                        for o, h in enumerate(handles):
                            if o == self.rank:
                                continue
                            q = C.c_void_p()
                            L.check(L.lib().td_peer_open(h, C.byref(q)), f"td_peer_open(rank {o})")
                            self.base[o] = int(q.value)
                            self._opened.append(int(q.value))
                except Exception as e:  # noqa: BLE001
                    err = f"rank {self.rank}: {e}"
            elif err is None:
                err = ""  # another rank failed to allocate; its message arrives below
            errs = [None] * self.world
            dist.all_gather_object(errs, err, group=group)
            errs = [e for e in errs if e]
            if errs or err is not None:
                self._release()
                raise PeerSetupError("peer data parallel set-up failed: " + "; ".join(errs))
        elif err is not None:
            raise PeerSetupError("peer data parallel set-up failed: " + err)
        self._bytes = torch.as_tensor(_DeviceBytes(self._local, lay.total_bytes, self), device=self.device)
        # local views
        self.flags = self._view(0, FLAG_BYTES, torch.int32)
        self.w1_bf16 = self._view(lay.off_w1, 2 * d * din, torch.bfloat16).view(d, din)
        self.w2_bf16 = self._view(lay.off_w2, 2 * d * d, torch.bfloat16).view(d, d)
        self.small_slots = self._view(lay.off_small, 4 * self.world * lay.small_numel, torch.float32).view(self.world, lay.small_numel)
        # host pointer arrays, built once
        arr = lambda ptrs: (C.c_void_p * len(ptrs))(*ptrs)  # noqa: E731
        self._flag_arrays = arr(self.base)  # every rank's flag block starts at offset 0
        self._dw_dst = {w: arr([self.base[o] + lay.grad_slot_offset(w, self.rank) for o in range(self.world)]) for w in (1, 2)}
        self._small_dst = arr([self.base[o] + lay.small_slot_offset(self.rank) for o in range(self.world)])
        self._w_dst = {w: arr([self.base[o] + lay.weight_rows_offset(w, self.rank) for o in range(self.world)]) for w in (1, 2)}
        self._folds = {}
        if self.world > 1:
            dist.barrier(group=group)  # every buffer is mapped everywhere before anyone stores into a peer

    def _view(self, off: int, nbytes: int, dtype):
        return self._bytes[off : off + nbytes].view(dtype)

    # -- destinations
    def dw_dst(self, which: int):
        """HOST array: where the rows owned by rank o of this rank's dW``which`` go (slot[rank] at rank o)."""
        return self._dw_dst[which]

    def grad_slots_ptr(self, which: int) -> int:
        """Device address of this rank's [world, rows, cols] slot block for weight ``which``."""
        return self._local + self.layout.grad_slot_offset(which, 0)

    # -- device operations (all asynchronous on torch's current stream)
    def signal(self, row: int):
        """+1 on counter ``row`` at every rank, after everything enqueued so far on the current stream."""
        L.launch_count += 1
        L.check(L.lib().td_peer_signal(self._flag_arrays, self.world, row * FLAG_ROW_STRIDE, L.stream_ptr()), "td_peer_signal")

    def wait(self, row: int, step: int):
        """The current stream waits until every rank has signalled ``row`` ``step`` times."""
        L.check(L.lib().td_peer_wait(C.c_void_p(self._local + self.layout.flag_offset(row)), 1, int(step) * self.world, self.timeout_s,
                                     L.stream_ptr()), "td_peer_wait")

    def fold_signal(self, row: int):
        """``PeerFold`` asking the last weight-gradient GEMM of a backward call for the +1 on counter ``row`` at every rank
        (instead of a ``signal(row)`` launch after it). Cached: the structure must stay alive until the call has returned."""
        f = self._folds.get(("signal", row))
        if f is None:
            f = L.PeerFold(C.cast(self._flag_arrays, C.c_void_p), row * FLAG_ROW_STRIDE, None, None, 0)
            self._folds[("signal", row)] = f
        return f

    def fold_post_small(self, small):
        """``PeerFold`` asking the backward's finisher launch to store ``small`` (fp32 [3 D], this rank's [db2 | dg | db1]) into this
        rank's slot at every rank as it writes it (instead of a ``post_small`` launch after it)."""
        if small.numel() != self.layout.small_numel:
            raise ValueError("small-vector size does not match the layout (3 D must be a multiple of 4)")
        return L.PeerFold(None, -1, C.cast(self._small_dst, C.c_void_p), C.c_void_p(small.data_ptr()), small.numel())

    def post_small(self, small):
        """``small`` fp32 [3 D] -> slot[rank] of every rank's small-vector block."""
        lay = self.layout
        if small.numel() != lay.small_numel:
            raise ValueError("small-vector size does not match the layout (3 D must be a multiple of 4)")
        L.launch_count += 1
        L.check(L.lib().td_peer_post(L.ptr(small), self._small_dst, self.world, lay.small_numel, L.stream_ptr()), "td_peer_post")

    def sum_small(self, out):
        lay = self.layout
        L.launch_count += 1
        L.check(L.lib().td_sum_slots(L.ptr(self.small_slots), lay.small_numel, self.world, L.ptr(out), lay.small_numel, L.stream_ptr()),
                "td_sum_slots")

    def adamw_rows(self, which: int, param_rows, exp_avg, exp_avg_sq, weight_decay, lr, betas, eps, step, grad_scale):
        """AdamW on this rank's row block of weight ``which`` from the summed slots; bf16 rows go to every rank."""
        lay = self.layout
        n = lay.slot1_numel if which == 1 else lay.slot2_numel
        if param_rows.numel() != n or not param_rows.is_contiguous():
            raise ValueError("param_rows must be this rank's contiguous row block")
        L.launch_count += 1
        L.check(L.lib().td_adamw_slots_step(L.ptr(param_rows), C.c_void_p(self.grad_slots_ptr(which)), n, self.world, L.ptr(exp_avg),
                                            L.ptr(exp_avg_sq), self._w_dst[which], self.world, n, weight_decay, lr, betas[0], betas[1],
                                            eps, int(step), grad_scale, None, L.stream_ptr()), "td_adamw_slots_step")

    def _release(self):
        """Undo a partial set-up (no collectives, no synchronisation)."""
        for q in self._opened:
            L.lib().td_peer_close(C.c_void_p(q))
        self._opened = []
        if self._local and self._torch_backing is None:
            L.lib().td_peer_free(C.c_void_p(self._local))
        self._local = 0

    def close(self):
        torch = self.torch
        if self._local:
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                import torch.distributed as dist

                dist.barrier(group=self.group)  # nobody still stores into a buffer that is about to be unmapped
            for q in self._opened:
                L.lib().td_peer_close(C.c_void_p(q))
            self._opened = []
            # the local allocation itself stays alive until the last aliasing tensor is gone (see _DeviceBytes / __del__)

    def __del__(self):
        try:
            if getattr(self, "_local", 0):
                if getattr(self, "_torch_backing", None) is None:
                    L.lib().td_peer_free(C.c_void_p(self._local))
                self._local = 0
        except Exception:
            pass
