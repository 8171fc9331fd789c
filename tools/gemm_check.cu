// Stand-alone bring-up / regression harness for tvt_gemm: compares against a CPU double-precision
// reference on bf16-rounded inputs and prints achieved TFLOP/s. Built by tools/build_tools.sh.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/tvt.h"

extern "C" void tvt_debug_set_mn_desc(unsigned lbo, unsigned sbo);
extern "C" void tvt_debug_set_epilogue(int mode);
extern "C" void tvt_debug_set_pair(int on);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

static float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
static uint32_t rng_state = 12345;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f; }

struct Case { int m, n, k; int amn, bmn, planes, splits; int epi; };  // epi: 0 plain, 1 bias+relu+residual (both with an fp32 copy: generic kernel), 2 atomic,
                                                                    //      3 bf16-only bias+residual, 4 bf16-only plain, 5 bf16-only bias+relu (the fast kernels)

static int run_case(const Case& c, bool verbose) {
  const int M = c.m, N = c.n, K = c.k;
  const long long lda = c.amn ? M : K, ldb = c.bmn ? N : K;
  const size_t na = (size_t)(c.amn ? K : M) * lda, nb = (size_t)(c.bmn ? K : N) * ldb;
  std::vector<float> A(na), B(nb), bias(N), res((size_t)M * N);
  for (auto& v : A) v = frand();
  for (auto& v : B) v = frand();
  for (auto& v : bias) v = frand();
  for (auto& v : res) v = bf16r(frand());
  std::vector<__nv_bfloat16> Ah(na), Al(na), Bh(nb), Bl(nb), Rh((size_t)M * N);
  for (size_t i = 0; i < na; ++i) { Ah[i] = __float2bfloat16_rn(A[i]); float h = __bfloat162float(Ah[i]); Al[i] = __float2bfloat16_rn(A[i] - h); if (c.planes == 1) A[i] = h; else A[i] = h + __bfloat162float(Al[i]); }
  for (size_t i = 0; i < nb; ++i) { Bh[i] = __float2bfloat16_rn(B[i]); float h = __bfloat162float(Bh[i]); Bl[i] = __float2bfloat16_rn(B[i] - h); if (c.planes == 1) B[i] = h; else B[i] = h + __bfloat162float(Bl[i]); }
  for (size_t i = 0; i < Rh.size(); ++i) Rh[i] = __float2bfloat16_rn(res[i]);
  __nv_bfloat16 *dA, *dAl, *dB, *dBl, *dR, *dO; float *dBias, *dF;
  CK(cudaMalloc(&dA, na * 2)); CK(cudaMalloc(&dAl, na * 2)); CK(cudaMalloc(&dB, nb * 2)); CK(cudaMalloc(&dBl, nb * 2));
  CK(cudaMalloc(&dR, (size_t)M * N * 2)); CK(cudaMalloc(&dO, (size_t)M * N * 2)); CK(cudaMalloc(&dBias, N * 4)); CK(cudaMalloc(&dF, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, Ah.data(), na * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAl, Al.data(), na * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bh.data(), nb * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBl, Bl.data(), nb * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dR, Rh.data(), (size_t)M * N * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBias, bias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dF, 0, (size_t)M * N * 4)); CK(cudaMemset(dO, 0, (size_t)M * N * 2));
  tvt_gemm_args g; memset(&g, 0, sizeof(g));
  g.a = dA; g.b = dB; if (c.planes == 2) { g.a_lo = dAl; g.b_lo = dBl; }
  g.m = M; g.n = N; g.k = K; g.lda = lda; g.ldb = ldb; g.a_mn_major = c.amn; g.b_mn_major = c.bmn; g.splits = c.splits; g.alpha = 1.0f;
  if (c.epi < 3) { g.out_f32 = dF; g.ld_f32 = N; } else { g.out_bf16 = dO; g.ld_bf16 = N; }
  if (c.epi == 3) { g.bias = dBias; g.residual = dR; g.residual_dtype = TVT_BF16; g.ld_residual = N; }
  if (c.epi == 5) { g.bias = dBias; g.act = TVT_ACT_RELU; }
  if (c.epi == 1) { g.bias = dBias; g.act = TVT_ACT_RELU; g.residual = dR; g.residual_dtype = TVT_BF16; g.ld_residual = N; g.out_bf16 = dO; g.ld_bf16 = N; }
  if (c.epi == 2) g.atomic_out = 1;
  int rc = tvt_gemm(&g, 0);
  if (rc != 0) { printf("tvt_gemm rc=%d: %s\n", rc, tvt_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<float> out((size_t)M * N);
  if (c.epi < 3) {
    CK(cudaMemcpy(out.data(), dF, out.size() * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<__nv_bfloat16> ob((size_t)M * N);
    CK(cudaMemcpy(ob.data(), dO, ob.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < ob.size(); ++i) out[i] = __bfloat162float(ob[i]);
  }
  // reference on a sample of rows (all rows when small)
  double max_err = 0, max_ref = 0; int bad = 0;
  const int row_step = M > 512 ? 37 : 1;
  for (int i = 0; i < M; i += row_step) {
    for (int j = 0; j < N; ++j) {
      double acc = 0;
      for (int k = 0; k < K; ++k) {
        const double a = c.amn ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k];
        const double b = c.bmn ? B[(size_t)k * ldb + j] : B[(size_t)j * ldb + k];
        acc += a * b;
      }
      if (c.epi == 1) { acc += bias[j]; if (acc < 0) acc = 0; acc += res[(size_t)i * N + j]; }
      if (c.epi == 3) acc += bias[j] + res[(size_t)i * N + j];
      if (c.epi == 5) { acc += bias[j]; if (acc < 0) acc = 0; }
      const double got = out[(size_t)i * N + j];
      const double err = fabs(got - acc);
      if (err > max_err) max_err = err;
      if (fabs(acc) > max_ref) max_ref = fabs(acc);
      const double tol = (c.planes == 2 ? 2e-4 : 2e-3) * sqrt((double)K) + 1e-3 + (c.epi >= 3 ? fabs(acc) / 128.0 : 0.0);
      if (!(err <= tol)) { if (bad < 3 && verbose) printf("   mismatch (%d,%d): got %g want %g\n", i, j, got, acc); ++bad; }
    }
  }
  printf("%s m=%d n=%d k=%d A:%s B:%s planes=%d splits=%d epi=%d  max_err=%.3e (max|ref|=%.2f) bad=%d\n", bad ? "FAIL" : "ok  ", M, N, K,
         c.amn ? "MN" : "K", c.bmn ? "MN" : "K", c.planes, c.splits, c.epi, max_err, max_ref, bad);
  cudaFree(dA); cudaFree(dAl); cudaFree(dB); cudaFree(dBl); cudaFree(dR); cudaFree(dO); cudaFree(dBias); cudaFree(dF);
  return bad ? 1 : 0;
}

static void bench(int M, int N, int K, int amn, int bmn, int planes, int splits, int epi = 0) {
  const long long lda = amn ? M : K, ldb = bmn ? N : K;
  const size_t na = (size_t)(amn ? K : M) * lda, nb = (size_t)(bmn ? K : N) * ldb;
  __nv_bfloat16 *dA, *dB, *dO; float* dF;
  CK(cudaMalloc(&dA, na * 2)); CK(cudaMalloc(&dB, nb * 2)); CK(cudaMalloc(&dO, (size_t)M * N * 2)); CK(cudaMalloc(&dF, (size_t)M * N * 4));
  CK(cudaMemset(dA, 0x3c, na * 2)); CK(cudaMemset(dB, 0x3c, nb * 2)); CK(cudaMemset(dF, 0, (size_t)M * N * 4));
  tvt_gemm_args g; memset(&g, 0, sizeof(g));
  g.a = dA; g.b = dB; if (planes == 2) { g.a_lo = dA; g.b_lo = dB; }
  g.m = M; g.n = N; g.k = K; g.lda = lda; g.ldb = ldb; g.a_mn_major = amn; g.b_mn_major = bmn; g.splits = splits; g.alpha = 1.0f;
  if (splits > 1) { g.out_f32 = dF; g.ld_f32 = N; g.atomic_out = 1; } else { g.out_bf16 = dO; g.ld_bf16 = N; }
  // epi: 1 = bias + relu + dropout, 2 = bias + bf16 residual, 3 = bf16 relu mask + dropout, 4 = bias + dropout + residual
  __nv_bfloat16* dR = nullptr; float* dBias = nullptr;
  if (epi) {
    CK(cudaMalloc(&dR, (size_t)M * N * 2)); CK(cudaMemset(dR, 0x3c, (size_t)M * N * 2));
    CK(cudaMalloc(&dBias, (size_t)N * 4)); CK(cudaMemset(dBias, 0, (size_t)N * 4));
    if (epi != 3) g.bias = dBias;
    if (epi == 1) g.act = TVT_ACT_RELU;
    if (epi == 1 || epi == 3 || epi == 4) { g.dropout_p = 0.5f; g.dropout_seed = 7; }
    if (epi == 2 || epi == 4) { g.residual = dR; g.ld_residual = N; g.residual_dtype = TVT_BF16; }
    if (epi == 3) { g.relu_mask = dR; g.ld_mask = N; g.mask_dtype = TVT_BF16; }
  }
  for (int i = 0; i < 3; ++i) tvt_gemm(&g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) tvt_gemm(&g, 0);
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  printf("bench m=%d n=%d k=%d A:%s B:%s planes=%d splits=%d epi=%d: %.3f ms  %.1f TFLOP/s\n", M, N, K, amn ? "MN" : "K", bmn ? "MN" : "K", planes, splits, epi, ms,
         2.0 * M * N * K * (planes == 2 ? 3 : 1) / ms * 1e-9);
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dF); cudaFree(dR); cudaFree(dBias);
}

int main(int argc, char** argv) {
  if (tvt_device_check() != 0) { printf("device check failed: %s\n", tvt_last_error()); return 1; }
  if (getenv("TVT_EPI_DBG")) tvt_debug_set_epilogue(atoi(getenv("TVT_EPI_DBG")));
  if (getenv("TVT_PAIR")) tvt_debug_set_pair(atoi(getenv("TVT_PAIR")));
  if (argc >= 5 && !strcmp(argv[1], "one")) {   // gemm_check one M N K [amn bmn planes splits epi]: a single shape, for ncu
    bench(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argc > 5 ? atoi(argv[5]) : 0, argc > 6 ? atoi(argv[6]) : 0,
          argc > 7 ? atoi(argv[7]) : 1, argc > 8 ? atoi(argv[8]) : 1, argc > 9 ? atoi(argv[9]) : 0);
    return 0;
  }
  int fails = 0;
  // K-major / K-major first (the forward product)
  Case kk[] = {{128, 128, 64, 0, 0, 1, 1, 0}, {128, 256, 128, 0, 0, 1, 1, 0}, {256, 512, 512, 0, 0, 1, 1, 0}, {200, 264, 200, 0, 0, 1, 1, 1},
               {2112, 512, 2048, 0, 0, 1, 1, 1}, {4096, 768, 768, 0, 0, 1, 1, 1}, {33024, 768, 768, 0, 0, 1, 1, 0}, {512, 512, 4096, 0, 0, 1, 8, 2},
               {300, 136, 328, 0, 0, 2, 1, 1}, {2112, 512, 512, 0, 0, 2, 1, 0}};
  for (auto& c : kk) fails += run_case(c, true);
  // dgrad (A K-major, B MN-major) and wgrad (both MN-major)
  Case mn[] = {{128, 128, 64, 0, 1, 1, 1, 0}, {128, 256, 128, 0, 1, 1, 1, 0}, {128, 128, 64, 1, 1, 1, 1, 0}, {256, 256, 256, 1, 1, 1, 1, 0}};
  int mn_fail = 0;
  for (auto& c : mn) mn_fail += run_case(c, true);
  if (mn_fail) {
    const unsigned cand[][2] = {{1024, 8192}, {8192, 128}, {128, 8192}, {1024, 1024}, {8192, 2048}};
    for (auto& cd : cand) {
      printf("--- retry MN-major with lbo=%u sbo=%u\n", cd[0], cd[1]);
      tvt_debug_set_mn_desc(cd[0], cd[1]);
      int f = 0;
      for (auto& c : mn) f += run_case(c, false);
      if (!f) { printf("+++ MN-major works with lbo=%u sbo=%u\n", cd[0], cd[1]); mn_fail = 0; break; }
    }
  }
  fails += mn_fail;
  Case more[] = {{2112, 2048, 512, 0, 1, 1, 1, 0}, {1000, 512, 776, 0, 1, 1, 1, 1}, {768, 3072, 33024, 1, 1, 1, 8, 2}, {512, 2048, 2112, 1, 1, 1, 4, 2},
                 {264, 136, 200, 1, 1, 1, 1, 0}, {512, 512, 2112, 1, 1, 2, 2, 2}, {2112, 512, 2048, 0, 1, 2, 1, 0}, {136, 896, 128, 0, 0, 1, 1, 1}};
  for (auto& c : more) fails += run_case(c, true);
  // the fast (bf16-only) kernels at sizes that take the CTA-pair path, ragged M included
  Case fastc[] = {{33024, 768, 768, 0, 0, 1, 1, 3}, {33024, 768, 2048, 0, 0, 1, 1, 3}, {33024, 768, 768, 0, 1, 1, 1, 3}, {33024, 2304, 768, 0, 0, 1, 1, 4},
                  {33160, 768, 512, 0, 0, 1, 1, 3}, {20000, 1024, 320, 0, 1, 1, 1, 5}, {4000, 512, 256, 0, 0, 1, 1, 5}, {3072, 768, 33024, 1, 1, 1, 2, 2}};
  for (auto& c : fastc) fails += run_case(c, true);
  printf("gemm_check: %d failing case(s)\n", fails);
  if (argc > 1 && !strcmp(argv[1], "bench")) {
    bench(33024, 768, 768, 0, 0, 1, 1);
    bench(33024, 2304, 768, 0, 0, 1, 1);
    bench(33024, 3072, 768, 0, 0, 1, 1);
    bench(33024, 768, 3072, 0, 0, 1, 1);
    bench(33024, 768, 3072, 0, 1, 1, 1);   // dgrad of linear1
    bench(33024, 3072, 768, 0, 1, 1, 1);   // dgrad of linear2
    bench(3072, 768, 33024, 1, 1, 1, 2);   // wgrad linear1
    bench(768, 3072, 33024, 1, 1, 1, 2);   // wgrad linear2
    bench(768, 768, 33024, 1, 1, 1, 8);    // wgrad out-proj
    bench(8192, 8192, 8192, 0, 0, 1, 1);
    bench(2112, 512, 2048, 0, 0, 1, 1);
    bench(33024, 768, 768, 0, 0, 2, 1);
  }
  return fails ? 1 : 0;
}
