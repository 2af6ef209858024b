// Host side of the tcgen05 GEMM: TMA tensor-map encoding (driver entry point resolved at run time, so the
// library links against cudart only) and the persistent launch.
#pragma once
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

#include "gemm_sm100.cuh"

namespace td {

// ---- error plumbing shared by the whole library (C-ABI returns codes, td_last_error() returns this text)
inline char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
#define TD_FAIL(code, ...)                                        \
  do {                                                            \
    snprintf(td::last_error_buf(), 512, __VA_ARGS__);             \
    return (code);                                                \
  } while (0)
#define TD_CUDA(expr)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) TD_FAIL(-100 - int(_e), "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

enum : int { TD_OK = 0, TD_ERR_ARG = -1, TD_ERR_UNSUPPORTED = -2, TD_ERR_DRIVER = -3 };

// ---- optional per-launch timing: when enabled (td_profile_enable) every kernel launch of the library is bracketed
// by a pair of CUDA events on the launch stream; td_profile_report() sums them per tag. Off by default (zero cost).
struct ProfRecord { const char* tag; cudaEvent_t e0, e1; double work; };
struct Profiler {
  bool on = false;
  std::vector<ProfRecord> recs;
  static Profiler& get() { static Profiler p; return p; }
};
struct ProfScope {
  cudaStream_t st; int idx = -1;
  // `work` = algorithmic FLOPs (GEMMs) or bytes (row kernels) of this launch, for the roofline report
  ProfScope(const char* tag, double work, cudaStream_t s) : st(s) {
    Profiler& p = Profiler::get();
    if (!p.on) return;
    ProfRecord r{tag, nullptr, nullptr, work};
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    p.recs.push_back(r);
    idx = int(p.recs.size()) - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(Profiler::get().recs[idx].e1, st); }
};

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  });
  return fn;
}

// bf16 operand with `rows` logical rows and contraction length K.
//   K-major : memory is [rows, K] (ld = elements between rows);  box = 64 (K) x box_rows
//   MN-major: memory is [K, rows] (ld = elements between k-rows); box = 64 (rows) x 64 (K)
inline int make_operand_map(CUtensorMap* map, const void* ptr, long long rows, long long K, long long ld,
                            bool mn_major, int box_rows) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) TD_FAIL(TD_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * 2) % 16)
    TD_FAIL(TD_ERR_ARG, "TMA operand needs a 16-byte aligned base and row pitch (ptr=%p ld=%lld)", ptr, ld);
  cuuint64_t gdim[2], gstride[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  if (!mn_major) {
    gdim[0] = (cuuint64_t)K; gdim[1] = (cuuint64_t)rows;
    box[0] = kBlockK; box[1] = (cuuint32_t)box_rows;
  } else {
    gdim[0] = (cuuint64_t)rows; gdim[1] = (cuuint64_t)K;
    box[0] = 64; box[1] = kBlockK;
  }
  gstride[0] = (cuuint64_t)ld * 2;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) TD_FAIL(TD_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return TD_OK;
}

inline int device_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
      cudaGetLastError();
      return 148;  // no device visible (size queries on a build box): assume a full B200
    }
  }
  return sms;
}

// Pool of {next unit, finished workers} counter pairs for the dynamic tile scheduler. Every launch takes the next
// pair; a pair is re-armed (zeroed) by the kernel that used it, so no per-launch memset is needed. 1024 pairs is far
// more than the launches that can be in flight at once.
inline int* next_sched_counter() {
  static int* pool = nullptr;
  static std::atomic<unsigned> seq{0};
  constexpr unsigned kPairs = 1024;
  if (!pool) {
    if (cudaMalloc(&pool, sizeof(int) * 2 * kPairs) != cudaSuccess) return nullptr;
    cudaMemset(pool, 0, sizeof(int) * 2 * kPairs);
  }
  return pool + 2 * (seq.fetch_add(1) % kPairs);
}

struct GemmOperand {
  const void* ptr;
  long long ld;   // elements
  bool mn_major;
};

// Choose a split-K factor that evens out the last wave: tall-K weight-gradient GEMMs have only a few hundred
// output tiles for 148 SMs. Only EPI_F32 may split (it reduces with red.add into a zeroed output).
inline int choose_splits(int num_tiles, int num_k_blocks, int workers, int max_splits) {
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= max_splits; ++s) {
    if (num_k_blocks / s < 8 && s > 1) break;
    const long long units = (long long)num_tiles * s;
    const long long waves = (units + workers - 1) / workers;
    const double eff = double(units) / double(waves * workers) - 0.01 * (s - 1);  // mild bias against splitting
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  return best;
}

template <int CTAS, bool A_MN, bool B_MN, int EPI>
int launch_gemm(GemmOperand a, GemmOperand b, typename ParamsFor<EPI>::type p, int splits, cudaStream_t stream,
                const char* tag = "gemm") {
  using S = GemmSmem<CTAS>;
  if (p.M <= 0 || p.N <= 0) return TD_OK;
  if (p.N % 32) TD_FAIL(TD_ERR_UNSUPPORTED, "GEMM N=%d must be a multiple of 32", p.N);
  if (p.K <= 0) TD_FAIL(TD_ERR_ARG, "GEMM K=%d must be positive (caller zero-fills empty contractions)", p.K);
  CUtensorMap ma, mb;
  int rc = make_operand_map(&ma, a.ptr, p.M, p.K, a.ld, A_MN, kBlockM);
  if (rc) return rc;
  rc = make_operand_map(&mb, b.ptr, p.N, p.K, b.ld, B_MN, kBlockN / CTAS);
  if (rc) return rc;
  p.num_m_blocks = (p.M + kBlockM * CTAS - 1) / (kBlockM * CTAS);
  p.num_n_blocks = (p.N + kBlockN - 1) / kBlockN;
  p.num_k_blocks = (p.K + kBlockK - 1) / kBlockK;
  const int workers = device_sm_count() / CTAS;
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  if (EPI != EPI_F32) splits = 1;
  if (splits <= 0) splits = choose_splits(tiles, p.num_k_blocks, workers, 4);
  if (splits > p.num_k_blocks) splits = p.num_k_blocks;
  p.k_blocks_per_split = (p.num_k_blocks + splits - 1) / splits;
  p.splits = (p.num_k_blocks + p.k_blocks_per_split - 1) / p.k_blocks_per_split;
  if (p.splits > 1) {
    // split-K reduces with red.add: the output must start from zero
    TD_CUDA(cudaMemsetAsync(p.out0, 0, sizeof(float) * (size_t)p.M * (size_t)p.ld_out, stream));
  }
  const long long units = (long long)tiles * p.splits;
  const int grid = int(units < workers ? units : workers) * CTAS;

  auto kern = gemm_bf16_kernel<CTAS, A_MN, B_MN, EPI>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    TD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  p.sched_counter = next_sched_counter();
  if (!p.sched_counter) TD_FAIL(TD_ERR_DRIVER, "cannot allocate the tile-scheduler counters");
  ProfScope prof(tag, 2.0 * double(p.M) * double(p.N) * double(p.K), stream);
  TD_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, p));
  return TD_OK;
}

}  // namespace td
