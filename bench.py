#!/usr/bin/env python
"""Benchmark of the ThinkDiff aligner training step (BASELINE.json metric: aligner train tokens/s).

    python bench.py [--gpus N] [--steps K] [--warmup W]                  # this repo's sm_100a path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # N > 1 (NCCL)
    python bench.py --impl reference ...                                  # the reference's CPU path (oracle port)

One step = pack (features + T5 targets, ragged -> cu_seqlens) -> aligner forward (bf16 autocast) -> masked MSE ->
aligner backward (-> gradient all-reduce over NCCL when N > 1) -> AdamW step, on one batch of synthetic Qwen2-VL-7B
features: BASELINE config 2 per GPU (64 sequences, valid length U{1..256}, 3584 -> 4096); N GPUs = 64 sequences per rank
(weak scaling; N = 8 is config 3's global batch of 512). A token = one valid (unpadded) row.

The JSON line: `value` = tokens/s with inputs resident in HBM (CUDA events, max over ranks); `e2e` = the same step fed
from pinned host buffers through the public API, H2D copies and a D2H read of the loss inside the timed region;
`roofline` = the dominant kernel (largest share of device time) from per-launch CUDA events in a separate profiled pass
of the same steps; `cpu_baseline` = the oracle's port of the reference module on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIN, D = 3584, 4096
SEQS_PER_GPU, MAX_LEN = 64, 256
FLOP_PER_TOKEN = 4 * DIN * D + 6 * D * D  # fwd 2 GEMMs + bwd 3 GEMMs (no dx), SURVEY.md section 8d
NUM_BATCHES = 4  # distinct input batches cycled through, so no step re-reads the previous step's inputs from L2
METRIC, UNIT = "aligner_train_tokens_per_sec", "tokens/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while a timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report that instead of guessing
            self.nv, self.err = None, repr(e)
        self.thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._stop.clear()
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *a):
        if self.thread is not None:
            self._stop.set()
            self.thread.join()

    def summary(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_run(steps: int, warmup: int, max_tokens: int = 2048):
    """The reference's CPU path: the oracle's port of build_vision_projector('mlp2x_gelu_t5_norm') (the reference is
    Python/PyTorch; its own modules cannot travel to the GPU box), fp32, all host threads, on a bounded sample of the
    config-2 batch (the first sequences up to `max_tokens` valid tokens): forward + MSE + backward + AdamW."""
    import torch

    from oracle import aligner_ref, pack_ref
    from thinkdiff_mlre_b200.train_step import make_reference_optimizer, synthetic_lvlm_batch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = synthetic_lvlm_batch(SEQS_PER_GPU, MAX_LEN, DIN, D, seed=1234, pin=False)
    lens, keep = b.lens.tolist(), 0
    while keep < len(lens) and sum(lens[: keep + 1]) <= max_tokens:
        keep += 1
    keep = max(keep, 1)
    bits, tbits = b.flat.view(torch.int16).numpy(), b.extras["flat_target"].view(torch.int16).numpy()
    x, _ = pack_ref.pack_from_flat(bits, b.src_row_start.tolist()[:keep], lens[:keep])      # the collater's job, on the CPU
    t, _ = pack_ref.pack_from_flat(tbits, b.src_row_start.tolist()[:keep], lens[:keep])
    x = torch.from_numpy(x).view(torch.bfloat16).float()
    t = torch.from_numpy(t).view(torch.bfloat16).float()
    m = aligner_ref.RefAligner(DIN, D)
    m.load_state_dict(aligner_ref.init_params_numpy(DIN, D, seed=0))
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.05)
    tokens = x.shape[0]

    def step():
        y = m(x)
        loss = torch.nn.functional.mse_loss(y, t)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": tokens * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {keep} sequences of the config-2 batch = {tokens} valid tokens, fp32, fwd+MSE+bwd+AdamW, {steps} steps after {warmup} warm-up",
            "ms_per_step": dt / steps * 1e3, "tokens_per_step": tokens}


def run_reference_arm(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 8)
    r = cpu_reference_run(steps, min(args.warmup, 2))
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 2), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(n):
    return {"workload": "ThinkDiff-LVLM aligner train step (BASELINE config 2 per GPU; N=8 = config 3 global batch 512)",
            "seqs_per_gpu": SEQS_PER_GPU, "global_batch": SEQS_PER_GPU * n, "max_len": MAX_LEN, "ragged": "len ~ U{1..256}",
            "din": DIN, "d": D, "loss": "masked_mse", "optimizer": "AdamW wd=0.05", "parallelism": f"dp{n}",
            "l2_policy": f"{NUM_BATCHES} distinct input batches cycled; per-step working set (~1.3 GB) exceeds the 126 MB L2"}


# ----------------------------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce after backward instead of overlapped")
    ap.add_argument("--no-pipeline", action="store_true", help="apply every parameter update inside its own step, on the compute stream")
    ap.add_argument("--peer", action="store_true", help="N > 1, EXPERIMENTAL: gradient exchange over NVLink peer memory from the GEMM epilogues, no NCCL kernels on the step (thinkdiff_mlre_b200/peer.py)")
    ap.add_argument("--no-shard", action="store_true", help="N > 1: all-reduce + replicated AdamW instead of reduce-scatter + row-sharded AdamW + all-gather")
    ap.add_argument("--profile-out", default="", help="write the per-kernel table (JSON) here")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="fused: this repo's one-pass AdamW (+ bf16 copies, per-bucket overlap); torch: torch.optim.AdamW(fused=True)")
    ap.add_argument("--loss-path", default="fused", choices=["fused", "module"],
                    help="fused: aligner.mse_loss_packed (y/dy stay on chip); module: forward() -> masked MSE -> backward()")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON); anything a library prints there meanwhile (e.g. NCCL's version
    # banner) is diverted to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        return run_reference_arm(args, emit)

    import torch
    import torch.distributed as dist

    import thinkdiff_mlre_b200 as td
    from thinkdiff_mlre_b200 import _lib as L
    from thinkdiff_mlre_b200.train_step import make_reference_optimizer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(args.gpus, 1) and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torch.distributed.run", file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    steps = args.steps

    torch.manual_seed(0)  # identical init on every rank (DDP broadcasts rank 0's; same seed is equivalent)
    aligner = td.ThinkDiffAligner(DIN, D).to(dev)
    if world > 1:
        sharded = args.optimizer == "fused" and args.loss_path == "fused" and not args.no_pipeline and not args.no_shard and D % world == 0
        aligner.enable_data_parallel(overlap=not args.no_overlap, defer_wait=args.optimizer == "fused", sharded=sharded,
                                     peer=bool(args.peer and sharded))
    opt = td.FusedAdamW(aligner, lr=1e-4, weight_decay=0.05) if args.optimizer == "fused" else make_reference_optimizer(aligner)
    pipelined = args.optimizer == "fused" and args.loss_path == "fused" and not args.no_pipeline
    stepper = td.AlignerTrainStep(aligner, opt, fused_loss=args.loss_path == "fused", pipelined=pipelined)

    host = [td.synthetic_lvlm_batch(SEQS_PER_GPU, MAX_LEN, DIN, D, seed=1234 + rank + 1000 * j) for j in range(NUM_BATCHES)]
    resident = [(b.flat.to(dev), b.src_row_start.to(dev), b.lens.to(dev), b.total_rows, b.l_max, b.extras["flat_target"].to(dev)) for b in host]
    tokens_per_step = [b.total_rows for b in host]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t)

    # ---- device-resident timing ------------------------------------------------------------------
    for i in range(warmup):
        stepper.step_device(*resident[i % NUM_BATCHES])
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.launch_count
    clocks = ClockSampler(local)
    with clocks:
        e0.record()
        t_host0 = time.perf_counter()
        for i in range(steps):
            loss = stepper.step_device(*resident[i % NUM_BATCHES])
        host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / steps  # CPU time to enqueue one step (no sync inside)
        stepper.flush()  # pipelined mode: the last step's updates are part of the timed work
        e1.record()
        sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = L.launch_count - launches0
    tokens = sum_over_ranks(float(sum(tokens_per_step[i % NUM_BATCHES] for i in range(steps))))
    value = tokens / (ms * 1e-3)
    final_loss = float(loss)

    # ---- end to end: pinned host buffers -> H2D -> step -> D2H loss, every step ---------------------
    e2e = None
    if not args.no_e2e:
        for i in range(3):
            float(stepper.step_host(host[i % NUM_BATCHES], dev))
        sync_all()
        e0.record()
        nxt = stepper.prefetch(host[0], dev)  # inside the timed region: every step's H2D is paid for
        for i in range(steps):
            cur = nxt
            if i + 1 < steps:
                nxt = stepper.prefetch(host[(i + 1) % NUM_BATCHES], dev)  # copy of step i+1 overlaps compute of step i
            loss_host = float(stepper.step_prefetched(cur))  # .item(): D2H read of the loss, as base_task.py:262
        stepper.flush()
        e1.record()
        sync_all()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        b0 = host[0]
        h2d = b0.flat.nbytes + b0.extras["flat_target"].nbytes + b0.src_row_start.nbytes + b0.lens.nbytes
        if os.environ.get("TD_E2E_AB"):  # developer A/B: the same loop with 1 / 2 / 4 copy streams
            for k in (1, 2, 4):
                stepper.copy_streams = k
                sync_all()
                e0.record()
                nxt = stepper.prefetch(host[0], dev)
                for i in range(steps):
                    cur = nxt
                    if i + 1 < steps:
                        nxt = stepper.prefetch(host[(i + 1) % NUM_BATCHES], dev)
                    float(stepper.step_prefetched(cur))
                stepper.flush()
                e1.record()
                sync_all()
                print(f"e2e A/B: {k} copy stream(s) {e0.elapsed_time(e1) / steps:.3f} ms/step (default run {ms_e2e / steps:.3f})", file=sys.stderr)
            stepper.copy_streams = 2
        e2e = {"value": tokens / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e / steps, "note": "every step: H2D of its flat bf16 features + T5 targets from pinned memory (row chunks on two copy streams, overlapping the previous step's compute) and loss.item()"}

    if os.environ.get("TD_HOST_PROFILE") and rank == 0:  # developer aid: where does the host time of a step go?
        import cProfile, io, pstats

        pr = cProfile.Profile()
        pr.enable()
        for i in range(100):
            stepper.step_device(*resident[i % NUM_BATCHES])
        pr.disable()
        stepper.flush()
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(22)
        print(buf.getvalue()[:5000], file=sys.stderr)
    elif os.environ.get("TD_HOST_PROFILE"):
        for i in range(100):
            stepper.step_device(*resident[i % NUM_BATCHES])
        stepper.flush()

    # ---- per-kernel device times (separate pass of the same steps; CUDA events around every launch) --
    psteps = min(steps, 20)
    prof_stepper, profile_pass = stepper, "same pipelined step (kernels of different streams overlap: per-launch times are upper bounds)"
    if world == 1 and pipelined:
        # one stream, no overlap between the AdamW side stream and the GEMMs: clean per-kernel times for the roofline
        stepper.flush()
        aligner._bwd_order, aligner._record_phase_events = "linear2_first", False
        prof_stepper = td.AlignerTrainStep(aligner, opt, fused_loss=True, pipelined=False)
        profile_pass = "sequential step: same kernels, same inputs, single stream (the timed region overlaps AdamW with the GEMMs)"
        for i in range(2):
            prof_stepper.step_device(*resident[i % NUM_BATCHES])
    L.profile_enable(True)
    for i in range(psteps):
        prof_stepper.step_device(*resident[i % NUM_BATCHES])
    prof_stepper.flush()
    torch.cuda.synchronize()
    prof = L.profile_report()
    L.profile_enable(False)
    peaks = measured_peaks()
    kernels = {}
    for tag, r in prof.items():
        per_launch_ms = r["ms"] / r["launches"]
        rate = r["work"] / (r["ms"] * 1e-3)
        if tag.startswith("gemm_"):
            kernels[tag] = {"launches_per_step": r["launches"] / psteps, "ms_per_launch": per_launch_ms, "bound": "tensor",
                            "achieved": rate / 1e12, "unit": "TFLOP/s", "frac": rate / 1e12 / peaks["bf16_tflops_sustained"]}
        else:
            kernels[tag] = {"launches_per_step": r["launches"] / psteps, "ms_per_launch": per_launch_ms, "bound": "hbm",
                            "achieved": rate / 1e9, "unit": "GB/s", "frac": rate / 1e9 / peaks["hbm_gbs"]}
    total_ms = sum(r["ms"] for r in prof.values()) / psteps
    for tag, r in prof.items():
        kernels[tag]["share_of_kernel_time"] = (r["ms"] / psteps) / total_ms
    dom = max(kernels, key=lambda k: kernels[k]["share_of_kernel_time"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu capture
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get(dom)
    roofline = {"kernel": dom, "bound": kernels[dom]["bound"], "achieved": kernels[dom]["achieved"],
                "peak": peaks["bf16_tflops_sustained"] if kernels[dom]["bound"] == "tensor" else peaks["hbm_gbs"],
                "unit": kernels[dom]["unit"], "frac": kernels[dom]["frac"], "traffic": traffic,
                "peak_source": f"MEASURED_PEAKS.json ({peaks['source']}); sustained bf16 figure: kernel timed inside a long step",
                "ms_per_launch": kernels[dom]["ms_per_launch"], "share_of_kernel_time": kernels[dom]["share_of_kernel_time"],
                "profile_pass": profile_pass}

    # the gradient all-reduce on its own (both buckets back to back, nothing else running): what overlap has to hide
    ar_alone = None
    if world > 1:
        from thinkdiff_mlre_b200.aligner import GradBuckets

        gb = GradBuckets(DIN, D, dev)
        gb.linear2.zero_(), gb.linear1.zero_()
        for _ in range(3):
            dist.all_reduce(gb.linear2), dist.all_reduce(gb.linear1)
        sync_all()
        e0.record()
        for _ in range(10):
            dist.all_reduce(gb.linear2), dist.all_reduce(gb.linear1)
        e1.record()
        sync_all()
        ar_alone = max_over_ranks(e0.elapsed_time(e1)) / 10

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(5, 2)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        step_tflops = value / world * FLOP_PER_TOKEN / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": dict(workload_config(world), loss_path=args.loss_path, optimizer_impl=args.optimizer, pipelined_updates=pipelined, sharded_optimizer=bool(world > 1 and aligner._dp.sharded), peer_exchange=bool(world > 1 and aligner._dp.peer)), "clocks": clocks.summary(), "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "step_tflops_per_gpu": step_tflops, "step_frac_of_bf16_peak": step_tflops / peaks["bf16_tflops_sustained"],
            "step_frac_of_nominal_2250": step_tflops / 2250.0, "tokens_per_step": tokens / steps, "final_loss": final_loss,
            "host_enqueue_ms_per_step": host_enqueue_ms, "allreduce_alone_ms": ar_alone, "kernels": kernels,
        }
        if args.profile_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
            with open(args.profile_out, "w") as f:
                json.dump({"kernels": kernels, "kernel_ms_per_step": total_ms, "ms_per_step": ms / steps, "peaks": peaks}, f, indent=1)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
