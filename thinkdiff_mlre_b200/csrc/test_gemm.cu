// Standalone bring-up / regression harness for the tcgen05 GEMM mainloop (not part of the shipped library).
// Compares every (CTA-pair?, A-major, B-major) variant with a naive CUDA-core GEMM on the device and times
// the aligner shapes.  Build: see Makefile target `test_gemm`.  Run on a B200:  ./test_gemm [quick]
#include <cmath>
#include <cstdlib>
#include <vector>

#include "gemm_host.cuh"

using namespace td;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e)); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__global__ void fill_bf16(__nv_bfloat16* p, long long n, uint32_t seed, float scale) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t x = uint32_t(i) * 2654435761u + seed;
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    float f = (float(x & 0xFFFF) / 65536.0f - 0.5f) * 2.0f * scale;
    p[i] = __float2bfloat16_rn(f);
  }
}

// D[m,n] = sum_k A(m,k) * B(n,k) with element accessors by major-ness
__global__ void ref_gemm(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int M, int N, int K, long long lda,
                         long long ldb, int a_mn, int b_mn) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float a = __bfloat162float(a_mn ? A[(long long)k * lda + m] : A[(long long)m * lda + k]);
    float b = __bfloat162float(b_mn ? B[(long long)k * ldb + n] : B[(long long)n * ldb + k]);
    acc = fmaf(a, b, acc);
  }
  D[(long long)m * N + n] = acc;
}

// ws == nullptr: leftover tiles run as a last partial wave; else the stream-K tail (see gemm_sm100.cuh)
template <int CTAS, bool A_MN, bool B_MN>
int run_f32(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int M, int N, int K, long long lda,
            long long ldb, void* ws, cudaStream_t st) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.out0 = D; p.ld_out = N; p.alpha = 1.0f;
  return launch_gemm<CTAS, A_MN, B_MN, EPI_F32>({A, lda, A_MN}, {B, ldb, B_MN}, p, ws, st);
}

typedef int (*RunFn)(const __nv_bfloat16*, const __nv_bfloat16*, float*, int, int, int, long long, long long, void*,
                     cudaStream_t);

struct Variant { const char* name; int ctas, a_mn, b_mn; RunFn fn; };

int main(int argc, char** argv) {
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  const int only = argc > 2 ? atoi(argv[2]) : -1;  // run a single variant (a trap kills the context)
  Variant variants[] = {
      {"cta1 A:K  B:K ", 1, 0, 0, run_f32<1, false, false>}, {"cta1 A:K  B:MN", 1, 0, 1, run_f32<1, false, true>},
      {"cta1 A:MN B:MN", 1, 1, 1, run_f32<1, true, true>},   {"cta2 A:K  B:K ", 2, 0, 0, run_f32<2, false, false>},
      {"cta2 A:K  B:MN", 2, 0, 1, run_f32<2, false, true>},  {"cta2 A:MN B:MN", 2, 1, 1, run_f32<2, true, true>},
  };
  // splits column: 0 = no stream-K workspace, 1 = with workspace (tail tiles cut along K)
  struct Shape { int M, N, K, splits; } shapes[] = {
      {128, 256, 64, 0}, {128, 256, 256, 1}, {256, 512, 512, 1}, {328, 768, 1096, 0}, {328, 768, 1096, 1}, {1000, 1024, 2048, 1},
      {512, 512, 4096, 1}, {4096, 4096, 2056, 0}, {4096, 4096, 2056, 1}, {4096, 3584, 1000, 1}, {2600, 4096, 1024, 1},
  };
  int fails = 0;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  void* sk_ws = nullptr;
  CK(cudaMalloc(&sk_ws, gemm_sk_workspace_bytes()));
  CK(cudaMemset(sk_ws, 0xFF, gemm_sk_workspace_bytes()));
  for (auto& sh : shapes) {
    const int M = sh.M, N = sh.N, K = sh.K;
    if (quick && (long long)M * N * K > (1ll << 31)) continue;
    __nv_bfloat16 *A, *B;
    float *D, *R;
    // allocate for either major-ness: [M,K] or [K,M] are the same element count
    CK(cudaMalloc(&A, sizeof(__nv_bfloat16) * (size_t)M * K));
    CK(cudaMalloc(&B, sizeof(__nv_bfloat16) * (size_t)N * K));
    CK(cudaMalloc(&D, sizeof(float) * (size_t)M * N));
    CK(cudaMalloc(&R, sizeof(float) * (size_t)M * N));
    fill_bf16<<<256, 256, 0, st>>>(A, (long long)M * K, 17u, 1.0f);
    fill_bf16<<<256, 256, 0, st>>>(B, (long long)N * K, 91u, 1.0f);
    std::vector<float> hd((size_t)M * N), hr((size_t)M * N);
    for (int vi = 0; vi < 6; ++vi) {
      if (only >= 0 && vi != only) continue;
      Variant& v = variants[vi];
      const long long lda = v.a_mn ? M : K, ldb = v.b_mn ? N : K;
      ref_gemm<<<dim3((N + 127) / 128, M), 128, 0, st>>>(A, B, R, M, N, K, lda, ldb, v.a_mn, v.b_mn);
      CK(cudaMemsetAsync(D, 0xFF, sizeof(float) * (size_t)M * N, st));  // NaN pattern: unwritten outputs show up
      int rc = v.fn(A, B, D, M, N, K, lda, ldb, sh.splits ? sk_ws : nullptr, st);
      if (rc) { printf("[%s] M=%d N=%d K=%d launch rc=%d: %s\n", v.name, M, N, K, rc, last_error_buf()); fails++; continue; }
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        printf("[%s] M=%d N=%d K=%d KERNEL FAILED: %s\n", v.name, M, N, K, cudaGetErrorString(e));
        return 3;  // context is dead after a trap
      }
      CK(cudaMemcpy(hd.data(), D, sizeof(float) * hd.size(), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hr.data(), R, sizeof(float) * hr.size(), cudaMemcpyDeviceToHost));
      double max_err = 0, max_ref = 0;
      long long bad = 0, first_bad = -1;
      for (size_t i = 0; i < hd.size(); ++i) {
        double d = fabs((double)hd[i] - (double)hr[i]);
        if (!(d <= 1e-3 * sqrt((double)K) + 1e-3 * fabs(hr[i]))) { bad++; if (first_bad < 0) first_bad = (long long)i; }
        if (d > max_err || d != d) max_err = d;
        if (fabs(hr[i]) > max_ref) max_ref = fabs(hr[i]);
      }
      printf("[%s] M=%5d N=%5d K=%5d streamK=%d  max_err=%.3e (max|ref|=%.2f) bad=%lld %s\n", v.name, M, N, K, sh.splits,
             max_err, max_ref, bad, bad ? "FAIL" : "ok");
      if (bad) {
        fails++;
        long long i = first_bad;
        printf("    first bad at (%lld,%lld): got %f want %f; D[0,0..3]= %f %f %f %f want %f %f %f %f\n", i / N, i % N,
               hd[i], hr[i], hd[0], hd[1], hd[2], hd[3], hr[0], hr[1], hr[2], hr[3]);
      }
    }
    cudaFree(A); cudaFree(B); cudaFree(D); cudaFree(R);
  }

  if (!quick) {
    // timing at the aligner shapes (cfg 2: M = 8224 tokens)
    struct Perf { const char* what; int M, N, K, variant; } perf[] = {
        {"fwd1  x.W1^T  pair", 8460, 4096, 3584, 3}, {"dh1   dh2.W2  pair", 8460, 4096, 4096, 4},
        {"dW2   dh2^T.h1 pair", 4096, 4096, 8460, 5}, {"dW1   dh0^T.x  pair", 4096, 3584, 8460, 5},
    };
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (auto& pf : perf) {
      if (only >= 0 && pf.variant != only) continue;
      Variant& v = variants[pf.variant];
      const int M = pf.M, N = pf.N, K = pf.K;
      __nv_bfloat16 *A, *B; float* D;
      CK(cudaMalloc(&A, sizeof(__nv_bfloat16) * (size_t)M * K));
      CK(cudaMalloc(&B, sizeof(__nv_bfloat16) * (size_t)N * K));
      CK(cudaMalloc(&D, sizeof(float) * (size_t)M * N));
      fill_bf16<<<256, 256, 0, st>>>(A, (long long)M * K, 17u, 1.0f);
      fill_bf16<<<256, 256, 0, st>>>(B, (long long)N * K, 91u, 1.0f);
      const long long lda = v.a_mn ? M : K, ldb = v.b_mn ? N : K;
      for (int splits = 0; splits <= 1; ++splits) {
        for (int i = 0; i < 3; ++i) v.fn(A, B, D, M, N, K, lda, ldb, splits ? sk_ws : nullptr, st);
        CK(cudaEventRecord(e0, st));
        const int iters = 20;
        for (int i = 0; i < iters; ++i) v.fn(A, B, D, M, N, K, lda, ldb, splits ? sk_ws : nullptr, st);
        CK(cudaEventRecord(e1, st));
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { printf("perf kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= iters;
        printf("perf %s M=%d N=%d K=%d streamK=%d: %.3f ms  %.1f TFLOP/s\n", pf.what, M, N, K, splits, ms,
               2.0 * M * N * K / ms * 1e-9);
      }
      cudaFree(A); cudaFree(B); cudaFree(D);
    }
  }
  printf(fails ? "RESULT: %d FAILURES\n" : "RESULT: ALL OK\n", fails);
  return fails ? 1 : 0;
}
