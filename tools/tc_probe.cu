// tc_probe.cu -- standalone check/timing of the tcgen05 3xTF32 GEMM (csrc/gemm_tc.cuh) against a double
// reference.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I closerlook3d_b200/csrc
//                    tools/tc_probe.cu -o gpurun_out/tc_probe
// Usage: tc_probe <case> <dbg_swap>     case 0: P x 144 x 80 (both k-fast)   1: P x 72 x 144 (B n-fast)
//                                        2: 80 x 144 x P split-K (both mn-fast) 3: small odd shape, scalar path
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#define TC_PROBE 1
#include "gemm_tc.cuh"
namespace cl3d {
void set_error(const char*) {}
}
using namespace cl3d;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
  const int cs = argc > 1 ? atoi(argv[1]) : 0, swap = argc > 2 ? atoi(argv[2]) : 0;
  const int P = 32768;
  int M, N, K, splits = 1;
  long long sa_m, sa_k, sb_k, sb_n;
  if (cs == 0) { M = P; N = 144; K = 80; sa_m = K; sa_k = 1; sb_k = 1; sb_n = K; }
  else if (cs == 1) { M = P; N = 72; K = 144; sa_m = K; sa_k = 1; sb_k = 80; sb_n = 1; }
  else if (cs == 2) { M = 80; N = 144; K = P; sa_m = 1; sa_k = 80; sb_k = 144; sb_n = 1; splits = 148; }
  else { M = 300; N = 37; K = 53; sa_m = 1; sa_k = 301; sb_k = 1; sb_n = 53; }
  size_t na = 0, nb = 0;
  na = (size_t)(M - 1) * sa_m + (size_t)(K - 1) * sa_k + 1;
  nb = (size_t)(K - 1) * sb_k + (size_t)(N - 1) * sb_n + 1;
  std::vector<float> ha(na), hb(nb);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
  for (auto& x : ha) x = rnd();
  for (auto& x : hb) x = rnd() * 0.3f;
  float *da, *db, *dc, *dp = nullptr;
  CK(cudaMalloc(&da, na * 4 + 64)); CK(cudaMalloc(&db, nb * 4 + 64)); CK(cudaMalloc(&dc, (size_t)M * N * 4));
  CK(cudaMemcpy(da, ha.data(), na * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), nb * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dc, 0xFF, (size_t)M * N * 4));
  int kps = (K + splits - 1) / splits; kps = (kps + 15) / 16 * 16; splits = (K + kps - 1) / kps;
  if (splits > 1) CK(cudaMalloc(&dp, (size_t)splits * M * N * 4));
  TcGemmArgs g{};
  g.A = da; g.sa_m = sa_m; g.sa_k = sa_k; g.B = db; g.sb_k = sb_k; g.sb_n = sb_n;
  g.M = M; g.N = N; g.K = K; g.k_per_split = kps; g.C = dc; g.sc_m = N; g.sc_n = 1; g.partial = dp;
  unsigned long long* dt; CK(cudaMalloc(&dt, 64 * 8)); CK(cudaMemset(dt, 0, 512));
  g.dbg_t = nullptr;
  tc_gemm_launch(g, splits, 0);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hc((size_t)M * N);
  if (dp) {
    std::vector<float> hp((size_t)splits * M * N);
    CK(cudaMemcpy(hp.data(), dp, hp.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t e = 0; e < hc.size(); ++e) { double a = 0; for (int z = 0; z < splits; ++z) a += hp[(size_t)z * M * N + e]; hc[e] = (float)a; }
  } else CK(cudaMemcpy(hc.data(), dc, hc.size() * 4, cudaMemcpyDeviceToHost));
  // reference on a sample of rows
  double maxerr = 0, maxref = 0; int nanc = 0;
  const int step = M > 2000 ? 97 : 1;
  for (int m = 0; m < M; m += step)
    for (int n = 0; n < N; ++n) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)ha[m * sa_m + k * sa_k] * (double)hb[k * sb_k + n * sb_n];
      const float c = hc[(size_t)m * N + n];
      if (!(c == c)) { ++nanc; continue; }
      maxerr = fmax(maxerr, fabs(r - c)); maxref = fmax(maxref, fabs(r));
    }
  printf("case %d swap %d: M=%d N=%d K=%d splits=%d  max|err|=%.3e  max|ref|=%.3e  rel=%.3e  nan=%d\n", cs, swap, M, N, K,
         splits, maxerr, maxref, maxerr / maxref, nanc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) tc_gemm_launch(g, splits, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < 50; ++i) tc_gemm_launch(g, splits, 0);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("  %.2f us per launch (warm, back-to-back)\n", ms * 1000.f / 50);
  g.dbg_t = dt;
  tc_gemm_launch(g, splits, 0); CK(cudaDeviceSynchronize());
  unsigned long long ht[64]; CK(cudaMemcpy(ht, dt, 512, cudaMemcpyDeviceToHost));
  printf("  stamps (ns since CTA start):");
  for (int i = 0; i < 16; ++i) if (ht[i]) printf(" [%d]%lld", i, (long long)(ht[i] - ht[0]));
  printf("\n  per step: entry wait store fence sync mma\n");
  for (int i = 0; i < 8; ++i) if (ht[16 + 6 * i]) { printf("   step %d:", i); for (int j = 0; j < 6; ++j) printf(" %lld", (long long)(ht[16 + 6 * i + j] - ht[0])); printf("\n"); }
  return 0;
}
