// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma (accumulators in TMEM)
// -> tcgen05.ld epilogue, with the aligner's elementwise / reduction work fused into the epilogue.
//
//   D[M,N] = A[M,K] * B[N,K]^T          (logical; fp32 accumulate)
//
// Each operand is either K-major (stored [rows, K], K contiguous -- activations in the forward pass, nn.Linear
// weights) or MN-major (stored [K, rows], rows contiguous -- what the backward pass has: dW = dY^T X contracts
// over the token dimension of two row-major activations, dX = dY W contracts over W's row index). The tensor
// core reads both through the 128-byte-swizzled canonical layouts, so no transposed copies are ever written.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue
// (TMEM lane quarter = warp % 4, two warps per quarter split the columns). CTAS=2 runs a CTA pair on one 256-row tile (tcgen05 cta_group::2): each CTA
// stages its own 128 rows of A and half of B, the leader issues the MMAs, both run their own epilogue.
//
// Reference semantics being fused (thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:58-63, under the bf16
// autocast of thinkdiff/tasks/base_task.py:237): Linear -> GELU(erf) -> Linear -> T5LayerNorm, each Linear/GELU
// output rounded to bf16, the norm statistics taken in fp32 from the bf16-rounded Linear2 output.
#pragma once
#include "ptx_sm100.cuh"

namespace td {

enum EpiKind : int {
  EPI_BF16 = 0,       // out0 = bf16(acc + bias?)
  EPI_BIAS_GELU = 1,  // out0 = h0 = bf16(acc + bias); out1 = bf16(gelu(h0))
  EPI_BIAS_SSQ = 2,   // out0 = h2 = bf16(acc + bias); red0[2 * n_blk + half][row] = sum over 128 cols of h2^2
  EPI_DGELU = 3,      // t = bf16(alpha * acc); out0 = bf16(t * gelu'(aux0)); red0[m_slab][col] = sum_rows out0
  EPI_F32 = 4,        // out0(fp32) = alpha * acc   (splits > 1: red.add into a zeroed out0)
  // Row-sharded output over peer memory (data-parallel weight gradients, reduce-scatter fused into the GEMM): output row r
  // belongs to rank o = r / scatter_rows and is written to scatter_dst[o] + (r - o * scatter_rows) * ld_out -- a buffer in
  // rank o's HBM mapped into this process (NVLink stores), or local memory for o == this rank. The accumulator chunk is
  // transposed in registers first so that every store instruction of a warp covers one full 128-byte line.
  EPI_F32_SCATTER = 5,
};
constexpr int kMaxPeers = 8;

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;  // in units of tile (BLOCK_M * CTAS, BLOCK_N, BLOCK_K)
  int splits, k_blocks_per_split;
  void* out0;
  void* out1;
  const void* aux0;
  const __nv_bfloat16* bias;
  float* red0;
  long long ld_out;  // elements
  float alpha;
  const float* alpha_ptr;  // optional device scalar multiplied into alpha (upstream loss gradient / GradScaler scale)
  int* sched_counter;      // [2] zero-initialised {next unit, finished workers}; re-armed by the kernel itself
};
// EPI_F32_SCATTER takes a larger parameter block; every other instantiation keeps the plain GemmParams signature
struct GemmScatterParams : GemmParams {
  float* scatter_dst[kMaxPeers];
  int scatter_rows;        // output rows per owner rank
  // Optional second problem in the same launch (the two weight-gradient GEMMs of a step share M = weight rows and K = tokens):
  // tiles [0, num_m_blocks * num_n_blocks) belong to problem 0, the rest to problem 1. One launch of 224 + 256 tiles fills the
  // 74 CTA pairs far more evenly (6.5 waves) than two un-split launches (3.03 and 3.46 waves).
  const CUtensorMap* maps2;  // DEVICE memory {A2, B2}; nullptr = single problem
  int N2, num_n_blocks2;
  long long ld_out2;
  const float* alpha_ptr2;
  float* scatter_dst2[kMaxPeers];
};
template <int EPI> struct ParamsFor { using type = GemmParams; };
template <> struct ParamsFor<EPI_F32_SCATTER> { using type = GemmScatterParams; };

constexpr int kBlockM = 128;  // rows per CTA (TMEM lanes)
constexpr int kBlockN = 256;  // accumulator columns (two stages fill the 512-column TMEM)
constexpr int kBlockK = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kAccStages = 2;

template <int CTAS>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;           // 16 KB
  static constexpr int kBBytes = (kBlockN / CTAS) * kBlockK * 2;  // 32 KB or 16 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (CTAS == 1) ? 4 : 7;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024;  // +1024 alignment slack
};

// ------------------------------------------------------------------------------------------ epilogues
// Eight epilogue warps: warp e (0..7) owns TMEM lane quarter (warp % 4) and column half e / 4 of the 256-column
// accumulator, i.e. 32 rows x 128 columns, walked in 4 chunks of 32 columns. Thread = one output row.
constexpr int kEpiWarps = 8;
constexpr int kEpiColsPerWarp = kBlockN / (kEpiWarps / 4);  // 128
constexpr int kEpiChunks = kEpiColsPerWarp / 32;            // 4

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

template <int EPI>
__device__ __forceinline__ void epilogue_tile(const typename ParamsFor<EPI>::type& p, uint32_t tmem_acc, int row0, int n0, int n_blk,
                                              int m_slab, int quarter, int half, int lane, int problem = 0) {
  const int row = row0 + quarter * 32 + lane;
  const bool row_ok = row < p.M;
  const int ncol0 = n0 + half * kEpiColsPerWarp;
  const uint32_t taddr = tmem_acc + (uint32_t(quarter * 32) << 16) + half * kEpiColsPerWarp;
  const long long row_off = (long long)row * p.ld_out;
  float ssq = 0.f;
  float alpha = p.alpha;
  if constexpr (EPI == EPI_F32 || EPI == EPI_DGELU) {
    if (p.alpha_ptr != nullptr) alpha *= __ldg(p.alpha_ptr);
  }
  int n_cols = p.N;
  if constexpr (EPI == EPI_F32_SCATTER) {
    const float* ap = problem ? p.alpha_ptr2 : p.alpha_ptr;
    if (ap != nullptr) alpha *= __ldg(ap);
    if (problem) n_cols = p.N2;
  }

  // software prefetch of the saved pre-activation (EPI_DGELU): chunk c+1 is in flight while chunk c is computed
  uint4 aux_next[4];
  auto load_aux = [&](int c) {
    const int col0 = ncol0 + c * 32;
    if (row_ok && col0 < p.N) {
      const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux0) + row_off + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) aux_next[q] = __ldg(ap + q);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) aux_next[q] = make_uint4(0, 0, 0, 0);
    }
  };
  if constexpr (EPI == EPI_DGELU) load_aux(0);

#pragma unroll 1
  for (int c = 0; c < kEpiChunks; ++c) {
    const int col0 = ncol0 + c * 32;
    if (col0 >= n_cols) break;  // N % 32 == 0 is enforced on the host
    uint32_t v[32];
    tmem_ld_32x32(taddr + c * 32, v);
    uint4 aux[4];
    if constexpr (EPI == EPI_DGELU) {
#pragma unroll
      for (int q = 0; q < 4; ++q) aux[q] = aux_next[q];
      if (c + 1 < kEpiChunks) load_aux(c + 1);
    }
    tmem_ld_wait();

    if constexpr (EPI == EPI_F32_SCATTER) {
      // 32 x 32 transpose across the warp (5 butterfly rounds of 16 exchanges): afterwards v[r] of lane j is element
      // (row r of this warp's 32-row slab, column col0 + j), so each store below writes 32 consecutive floats
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if ((i & off) == 0) {
            const uint32_t send = upper ? v[i] : v[i + off];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, off);
            if (upper) v[i] = recv; else v[i + off] = recv;
          }
        }
      }
      const int slab_row0 = row0 + quarter * 32;
      const long long ld = problem ? p.ld_out2 : p.ld_out;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const int orow = slab_row0 + r;  // warp-uniform
        if (orow < p.M) {
          const int owner = orow / p.scatter_rows;
          // select with constant indices: no local copy of the parameter struct
          float* base = problem ? p.scatter_dst2[0] : p.scatter_dst[0];
#pragma unroll
          for (int o = 1; o < kMaxPeers; ++o) base = (owner == o) ? (problem ? p.scatter_dst2[o] : p.scatter_dst[o]) : base;
          base[(long long)(orow - owner * p.scatter_rows) * ld + col0 + lane] = alpha * __uint_as_float(v[r]);
        }
      }
    } else if constexpr (EPI == EPI_F32) {
      float* out = reinterpret_cast<float*>(p.out0) + row_off + col0;
      if (row_ok) {
        if (p.splits == 1) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(alpha * __uint_as_float(v[j]), alpha * __uint_as_float(v[j + 1]),
                                   alpha * __uint_as_float(v[j + 2]), alpha * __uint_as_float(v[j + 3]));
            *reinterpret_cast<float4*>(out + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + j),
                         "f"(alpha * __uint_as_float(v[j])), "f"(alpha * __uint_as_float(v[j + 1])),
                         "f"(alpha * __uint_as_float(v[j + 2])), "f"(alpha * __uint_as_float(v[j + 3]))
                         : "memory");
          }
        }
      }
    } else if constexpr (EPI == EPI_DGELU) {
      // dh0 = bf16( bf16(dh1) * gelu'(h0) ): dh1 is rounded to bf16 first, as autograd materialises it.
      float colsum[32];
      uint4* o0 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out0) + row_off + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float x[8], h[8];
        unpack8(aux[q], x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = bf16_round(alpha * __uint_as_float(v[q * 8 + j]));
          h[j] = row_ok ? bf16_round(t * gelu_grad_fast(x[j])) : 0.f;
          colsum[q * 8 + j] = h[j];
        }
        if (row_ok) o0[q] = pack8(h);
      }
      // Column sums over this warp's 32 rows by a butterfly transpose-reduce: after the 5 rounds lane j holds
      // the sum over lanes of colsum[j]. 31 shuffles for 32 columns.
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
          const float send = upper ? colsum[i] : colsum[i + off];
          const float keep = upper ? colsum[i + off] : colsum[i];
          colsum[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      p.red0[(long long)(m_slab * 4 + quarter) * p.N + col0 + lane] = colsum[0];
    } else {
      // bias (bf16, same 32 columns for every thread of the warp -> broadcast loads)
      uint4* o0 = p.out0 ? reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out0) + row_off + col0) : nullptr;
      uint4* o1 = nullptr;
      if constexpr (EPI == EPI_BIAS_GELU) o1 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out1) + row_off + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float b[8], h[8];
        if (p.bias != nullptr) {
          unpack8(__ldg(reinterpret_cast<const uint4*>(p.bias + col0) + q), b);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) b[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = bf16_round(__uint_as_float(v[q * 8 + j]) + b[j]);
        if (row_ok && o0 != nullptr) o0[q] = pack8(h);  // inference skips saving the pre-activation
        if constexpr (EPI == EPI_BIAS_GELU) {
          float g[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = gelu_fast(h[j]);
          if (row_ok) o1[q] = pack8(g);
        }
        if constexpr (EPI == EPI_BIAS_SSQ) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ssq = fmaf(h[j], h[j], ssq);
        }
      }
    }
  }
  if constexpr (EPI == EPI_BIAS_SSQ) {
    if (row_ok) p.red0[(long long)(n_blk * 2 + half) * p.M + row] = ssq;
  }
}

// ------------------------------------------------------------------------------------------ kernel
// Work distribution: a unit = (output tile, split-K slice). The first unit of every CTA (pair) is its own index; after
// that the leader's producer thread claims units from a global atomic counter and publishes each claim through a small
// shared-memory ring (`sched_*`) to the MMA thread and the epilogue warps -- and, for a pair, to the peer CTA over
// DSMEM. CTAs that become resident late (e.g. because an NCCL kernel holds some SMs) simply find less work left,
// instead of delaying a statically assigned share of the tiles.
constexpr int kSchedStages = 4;

template <int CTAS, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const typename ParamsFor<EPI>::type p) {
  using S = GemmSmem<CTAS>;
  constexpr int kStages = S::kStages;
  constexpr int kBRows = kBlockN / CTAS;  // B rows staged by one CTA

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full_bar = bars;                                  // [kStages]   TMA -> MMA
  uint64_t* empty_bar = full_bar + kStages;                   // [kStages]   MMA -> TMA
  uint64_t* acc_full_bar = empty_bar + kStages;               // [kAccStages] MMA -> epilogue
  uint64_t* acc_empty_bar = acc_full_bar + kAccStages;        // [kAccStages] epilogue -> MMA
  uint64_t* sched_full_bar = acc_empty_bar + kAccStages;      // [kSchedStages] scheduler -> consumers
  uint64_t* sched_empty_bar = sched_full_bar + kSchedStages;  // [kSchedStages] consumers -> scheduler (leader's copy is used)
  int* sched_unit = reinterpret_cast<int*>(sched_empty_bar + kSchedStages);  // [kSchedStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_unit + kSchedStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  // consumers of a published unit, per CTA: 8 epilogue warps + 1 (the MMA thread in the leader, the producer in the peer)
  constexpr uint32_t kSchedConsumers = (kEpiWarps + 1) * CTAS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&acc_full_bar[s], 1);
      mbar_init(&acc_empty_bar[s], kEpiWarps * CTAS);
    }
    for (int s = 0; s < kSchedStages; ++s) {
      mbar_init(&sched_full_bar[s], 1);
      mbar_init(&sched_empty_bar[s], kSchedConsumers);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<CTAS>(tmem_slot, 512);
    tmem_relinquish<CTAS>();
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_workers = gridDim.x / CTAS;
  const int worker = blockIdx.x / CTAS;
  const int tiles0 = p.num_m_blocks * p.num_n_blocks;
  int tiles_all = tiles0;
  if constexpr (EPI == EPI_F32_SCATTER) {
    if (p.maps2 != nullptr) tiles_all += p.num_m_blocks * p.num_n_blocks2;
  }
  const int num_tiles = tiles_all;
  const int num_units = num_tiles * p.splits;
  constexpr int kGroupM = 8;  // rasterise m-blocks in groups so that concurrently running tiles share operands in L2

  // returns the problem index (always 0 unless this is a grouped EPI_F32_SCATTER launch)
  auto unit_coords = [&](int unit, int& m_blk, int& n_blk, int& kb0, int& kb1) -> int {
    int tile = unit % num_tiles;
    const int split = unit / num_tiles;
    int nnb = p.num_n_blocks;
    int problem = 0;
    if constexpr (EPI == EPI_F32_SCATTER) {
      if (tile >= tiles0) {
        problem = 1;
        tile -= tiles0;
        nnb = p.num_n_blocks2;
      }
    }
    const int group = tile / (kGroupM * nnb);
    const int first_m = group * kGroupM;
    const int gsize = min(kGroupM, p.num_m_blocks - first_m);
    const int in_group = tile - group * kGroupM * nnb;
    m_blk = first_m + in_group % gsize;
    n_blk = in_group / gsize;
    kb0 = split * p.k_blocks_per_split;
    kb1 = min(p.num_k_blocks, kb0 + p.k_blocks_per_split);
    return problem;
  };
  // consumer side of the scheduler ring: wait for slot `ss`, read the unit, release the slot (one arrive per warp)
  auto sched_consume = [&](int ss, uint32_t sphase, bool whole_warp) -> int {
    mbar_wait_cluster(&sched_full_bar[ss], sphase);
    const int unit = *reinterpret_cast<volatile int*>(&sched_unit[ss]);
    if (whole_warp) __syncwarp();
    if (!whole_warp || lane == 0) {
      if constexpr (CTAS == 1) mbar_arrive(&sched_empty_bar[ss]);
      else mbar_arrive_cluster(&sched_empty_bar[ss], 0);
    }
    return unit;
  };

  if (warp == 0) {
    // ===================================================== scheduler + TMA producer
    if (lane == 0) {
      int stage = 0, ss = 0;
      uint32_t phase = 0, sphase = 0;
      if constexpr (EPI == EPI_F32_SCATTER) {
        // tensor maps in global memory (written by a copy before this launch) need an acquire through the tensormap proxy
        if (p.maps2 != nullptr) {
          fence_tensormap_acquire(p.maps2);
          fence_tensormap_acquire(p.maps2 + 1);
        }
      }
      // the claim for the NEXT unit is issued before the current unit's loads, so the atomic's round trip is hidden
      int claimed = worker;
      while (true) {
        int unit;
        if (leader) {
          unit = claimed;
          if (unit < num_units) claimed = atomicAdd(p.sched_counter, 1) + num_workers;
          if (unit >= num_units) unit = -1;
          mbar_wait_cluster(&sched_empty_bar[ss], sphase ^ 1);  // every consumer (both CTAs) has read the old value
          sched_unit[ss] = unit;
          mbar_arrive(&sched_full_bar[ss]);  // release.cta: orders the store above for this CTA's consumers
          if constexpr (CTAS == 2) {
            st_shared_cluster_s32(&sched_unit[ss], 1, unit);
            mbar_arrive_cluster(&sched_full_bar[ss], 1);  // release.cluster: orders the remote store
          }
        } else {
          unit = sched_consume(ss, sphase, false);
        }
        if (++ss == kSchedStages) { ss = 0; sphase ^= 1; }
        if (unit < 0) break;
        int m_blk, n_blk, kb0, kb1;
        const int problem = unit_coords(unit, m_blk, n_blk, kb0, kb1);
        const CUtensorMap* map_a = &tmap_a;
        const CUtensorMap* map_b = &tmap_b;
        if constexpr (EPI == EPI_F32_SCATTER) {
          if (problem) {
            map_a = p.maps2;
            map_b = p.maps2 + 1;
          }
        }
        const int a_row0 = m_blk * (kBlockM * CTAS) + cta_rank * kBlockM;
        const int b_row0 = n_blk * kBlockN + cta_rank * kBRows;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes * CTAS);
          const int k0 = kb * kBlockK;
          if constexpr (!A_MN) {
            if constexpr (CTAS == 1) tma_load_2d(sa, map_a, &full_bar[stage], k0, a_row0);
            else tma_load_2d_pair(sa, map_a, &full_bar[stage], k0, a_row0);
          } else {
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j) {
              if constexpr (CTAS == 1) tma_load_2d(sa + j * (kBlockK * 128), map_a, &full_bar[stage], a_row0 + j * 64, k0);
              else tma_load_2d_pair(sa + j * (kBlockK * 128), map_a, &full_bar[stage], a_row0 + j * 64, k0);
            }
          }
          if constexpr (!B_MN) {
            if constexpr (CTAS == 1) tma_load_2d(sb, map_b, &full_bar[stage], k0, b_row0);
            else tma_load_2d_pair(sb, map_b, &full_bar[stage], k0, b_row0);
          } else {
#pragma unroll
            for (int j = 0; j < kBRows / 64; ++j) {
              if constexpr (CTAS == 1) tma_load_2d(sb + j * (kBlockK * 128), map_b, &full_bar[stage], b_row0 + j * 64, k0);
              else tma_load_2d_pair(sb + j * (kBlockK * 128), map_b, &full_bar[stage], b_row0 + j * 64, k0);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (leader) {
        // the last worker to run out of units re-arms the counter pair for the next launch that uses this slot
        __threadfence();
        if (atomicAdd(p.sched_counter + 1, 1) == num_workers - 1) {
          p.sched_counter[0] = 0;
          p.sched_counter[1] = 0;
          __threadfence();
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM * CTAS, kBlockN, A_MN, B_MN);
      // K-major SW128: rows 128 B apart, 8-row groups 1024 B apart (SBO); LBO unused.
      // MN-major SW128: 64-element row chunks; next 8 k-rows 1024 B (SBO), next 64 MN elements one box (LBO).
      constexpr uint32_t a_lbo = A_MN ? kBlockK * 128 : 16, b_lbo = B_MN ? kBlockK * 128 : 16;
      constexpr uint32_t sbo = 1024;
      constexpr uint32_t a_kstep = A_MN ? kUmmaK * 128 : kUmmaK * 2;  // bytes per UMMA_K step
      constexpr uint32_t b_kstep = B_MN ? kUmmaK * 128 : kUmmaK * 2;
      int stage = 0, ss = 0, acc = 0;
      uint32_t phase = 0, sphase = 0, acc_phase = 0;
      while (true) {
        const int unit = sched_consume(ss, sphase, false);
        if (++ss == kSchedStages) { ss = 0; sphase ^= 1; }
        if (unit < 0) break;
        int m_blk, n_blk, kb0, kb1;
        unit_coords(unit, m_blk, n_blk, kb0, kb1);
        mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kBlockN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(sa + k * a_kstep, a_lbo, sbo);
            const uint64_t bdesc = make_smem_desc_sw128(sb + k * b_kstep, b_lbo, sbo);
            umma_bf16<CTAS>(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit<CTAS>(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit<CTAS>(&acc_full_bar[acc]);  // accumulator complete -> epilogue(s)
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue warps (TMEM lane quarter = warp % 4)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0, ss = 0;
    uint32_t acc_phase = 0, sphase = 0;
    while (true) {
      const int unit = sched_consume(ss, sphase, true);
      if (++ss == kSchedStages) { ss = 0; sphase ^= 1; }
      if (unit < 0) break;
      int m_blk, n_blk, kb0, kb1;
      const int problem = unit_coords(unit, m_blk, n_blk, kb0, kb1);
      mbar_wait(&acc_full_bar[acc], acc_phase);
      tc_fence_after();
      const int m_slab = m_blk * CTAS + cta_rank;
      if (kb1 > kb0)
        epilogue_tile<EPI>(p, tmem_base + acc * kBlockN, m_slab * kBlockM, n_blk * kBlockN, n_blk, m_slab, quarter, half, lane,
                           problem);
      tc_fence_before();
      __syncwarp();  // all 32 lanes have drained their TMEM loads; one (release) arrive per warp
      if (lane == 0) {
        if constexpr (CTAS == 1) mbar_arrive(&acc_empty_bar[acc]);
        else mbar_arrive_cluster(&acc_empty_bar[acc], 0);
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  // teardown: everyone (both CTAs of a pair) must be done with TMEM and the barriers before it is freed
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CTAS>(tmem_base, 512);
  }
}

}  // namespace td
