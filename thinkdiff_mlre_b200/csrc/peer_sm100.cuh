// Data-parallel exchange over NVLink peer memory (no NCCL kernels on the step): the kernels that go with the
// EPI_F32_SCATTER GEMM epilogue. Every rank owns the rows [r * rows/N, (r + 1) * rows/N) of each weight matrix:
//
//   weight-gradient GEMM (rank s)  --stores-->  slot[s] of the owner's gradient buffer     (gemm_sm100.cuh)
//   peer_signal_kernel             --count--->  +1 on the owner's counter                   (after the GEMM, same stream)
//   td_peer_wait (owner)           the stream waits until the counter reaches t * N (stream memory op; or peer_wait_kernel)
//   adamw_slots_kernel (owner)     g = slot[0] + slot[1] + ... (fixed order: bit-reproducible, identical on every rank),
//                                  AdamW on the owner's fp32 master rows, bf16 rows stored into EVERY rank's compute copy
//   peer_signal_kernel             "owner o has written the weights of step t"; the next forward waits on those flags
//
// Flags are monotonically increasing counters in the destination rank's memory: release/acquire at system scope, no resets, no
// barriers.
//
// Reference being replaced: DDP's bucketed NCCL all-reduce of the aligner gradients + the replicated optimizer step
// (thinkdiff/runners/runner_base.py:88-92, :98-127; thinkdiff/tasks/base_task.py:247-258).
#pragma once
#include <cstdio>

#include "gemm_sm100.cuh"
#include "rowops_sm100.cuh"

namespace td {

struct PeerPtrs {
  void* p[kMaxPeers];
};

__device__ __forceinline__ void* select_peer(const PeerPtrs& a, int i) {
  void* r = a.p[0];  // constant indices only: the parameter struct stays in the constant bank
#pragma unroll
  for (int o = 1; o < kMaxPeers; ++o) r = (i == o) ? a.p[o] : r;
  return r;
}

// Thread i < n adds 1 to counter `slot` of the int32 flag array that lives in rank i's memory (release at system scope: everything
// this stream did before -- the GEMM's stores into that rank's slots, the AdamW's stores into its weight copy -- is visible to
// whoever observes the new count). A row's counter reaches t * world when every rank has signalled step t.
__global__ void peer_signal_kernel(const PeerPtrs flags, int n, int slot) {
  const int i = threadIdx.x;
  if (i < n) {
    int* f = static_cast<int*>(select_peer(flags, i)) + slot;
    __threadfence_system();
    asm volatile("red.release.sys.global.add.s32 [%0], 1;" ::"l"(f) : "memory");
  }
}

// Thread i < n spins until flags[i] >= value (local memory, written by the peers' peer_signal_kernel). The poll is a RELAXED
// load with a pause in between -- an acquire at system scope on every iteration is a fence storm that slows every other kernel
// on the GPU (measured: all GEMMs 1.5x slower while a waiter spun) -- and the acquire fence is executed once, after the flag
// has flipped.
__global__ void peer_wait_kernel(const int* flags, int n, int value, unsigned long long timeout_ns) {
  const int i = threadIdx.x;
  if (i < n) {
    const uint64_t t0 = global_timer_ns();
    unsigned spins = 0;
    while (true) {
      int v;
      asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
      if (v >= value) break;
      __nanosleep(400);
      if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) {
        printf("thinkdiff_b200: peer wait timed out (flag %d of %d is %d, expected >= %d)\n", i, n, v, value);
        __trap();
      }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
}

// Copy n4 float4 from `src` into dst.p[0..n_dst) (the small-vector gradients go to every rank's slot for this rank).
__global__ void __launch_bounds__(256) peer_post_kernel(const float4* __restrict__ src, const PeerPtrs dst, int n_dst, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(src + i);
#pragma unroll
    for (int o = 0; o < kMaxPeers; ++o)
      if (o < n_dst) static_cast<float4*>(dst.p[o])[i] = v;
  }
}

// out[i] = slots[0][i] + slots[1][i] + ... + slots[n-1][i]  (slot s at slots + s * stride4), in that order on every rank.
__global__ void __launch_bounds__(256) sum_slots_kernel(const float4* __restrict__ slots, long long stride4, int n_slots,
                                                        float4* __restrict__ out, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = __ldg(slots + i);
    for (int s = 1; s < n_slots; ++s) {
      const float4 b = __ldg(slots + (long long)s * stride4 + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    out[i] = a;
  }
}

// AdamW on this rank's rows of one weight matrix. The gradient is the sum of the N per-source-rank slots (each already
// scaled by 1/N in its GEMM epilogue); the updated rows are rounded to bf16 and stored into every rank's compute copy
// (p_bf16.p[o] = address of the same row block inside rank o's bf16 weight). Same update rule as adamw_kernel.
struct AdamSlotsParams {
  float* p; float* m; float* v;
  const float* slots; long long slot_stride;  // elements between slots
  int n_slots;
  PeerPtrs p_bf16; int n_dst;
  long long n;
  float weight_decay, lr, beta1, beta2, eps, bias_c1, sqrt_bias_c2, grad_scale;
  const StepCtl* ctl;  // optional device-side step control (see rowops_sm100.cuh)
};

__global__ void __launch_bounds__(256) adamw_slots_kernel(const AdamSlotsParams a) {
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long slot4 = a.slot_stride >> 2;
  float bias_c1 = a.bias_c1, sqrt_bias_c2 = a.sqrt_bias_c2, grad_scale = a.grad_scale;
  if (a.ctl != nullptr) {
    if (a.ctl->skip) return;  // skipped step: every rank's bf16 rows stay as they are
    bias_c1 = a.ctl->bias_c1; sqrt_bias_c2 = a.ctl->sqrt_bias_c2; grad_scale *= a.ctl->grad_mult;
  }
  const float decay = 1.0f - a.lr * a.weight_decay;
  const float step_size = a.lr / bias_c1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float4 g4 = __ldg(reinterpret_cast<const float4*>(a.slots) + i);
    for (int s = 1; s < a.n_slots; ++s) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(a.slots) + (long long)s * slot4 + i);
      g4.x += b.x; g4.y += b.y; g4.z += b.z; g4.w += b.w;
    }
    float pp[4] = {p.x, p.y, p.z, p.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w}, mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float g = gg[q] * grad_scale;
      pp[q] *= decay;
      mm[q] = a.beta1 * mm[q] + (1.0f - a.beta1) * g;
      vv[q] = a.beta2 * vv[q] + (1.0f - a.beta2) * g * g;
      const float denom = sqrtf(vv[q]) / sqrt_bias_c2 + a.eps;
      pp[q] -= step_size * (mm[q] / denom);
    }
    reinterpret_cast<float4*>(a.p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(a.m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(a.v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    const uint2 w = make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
#pragma unroll
    for (int o = 0; o < kMaxPeers; ++o)
      if (o < a.n_dst) static_cast<uint2*>(a.p_bf16.p[o])[i] = w;
  }
}

}  // namespace td
