// gemm.cu -- fp32 FMA GEMM with arbitrary element strides and optional split-K (sm_100a).
//
// Used for the per-point products of the fused PointWiseMLP (see pwmlp.cu): the reference's per-neighbour
// 1x1 conv over [dp; f_i; f_j - f_i] (/root/reference/pytorch/models/local_aggregation_operators.py:254-257,
// 288-295) is refactored into  A = f (Wc - Wr)^T  and  Bv = f Wr^T  per POINT, i.e. (B*N x C) x (C x 2*Cout)
// products: K*C is far too small to fill a tcgen05 tile and TF32/BF16 operands would break the fp32 1e-5
// parity bar, so these stay vectorised fp32 FMA (BASELINE.json north_star).
//
//   C[m][n] = sum_k A[m*sa_m + k*sa_k] * B[k*sb_k + n*sb_n]
// 64x64 output tile, 16-deep k tiles, 256 threads, 4x4 outputs per thread.  With splitk > 1 the k range is
// divided among gridDim.z CTAs that write partial tiles, reduced afterwards in a fixed order.
#include "common.cuh"

namespace cl3d {

constexpr int kGM = 64, kGN = 64, kGK = 16;

__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float* __restrict__ A, long long sa_m, long long sa_k,
                                                            const float* __restrict__ B, long long sb_k,
                                                            long long sb_n, int M, int N, int K, int k_per_split,
                                                            float* __restrict__ C, long long ldc,
                                                            float* __restrict__ partial) {
  __shared__ float sA[kGK][kGM + 4];
  __shared__ float sB[kGK][kGN + 4];
  const int m0 = blockIdx.y * kGM, n0 = blockIdx.x * kGN;
  const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, each a 4x4 block
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping chosen at run time so that the fastest-varying thread index follows the unit stride
  const bool a_k_fast = (sa_k == 1), b_n_fast = (sb_n == 1);
  for (int k0 = kbeg; k0 < kend; k0 += kGK) {
    // A tile: kGM x kGK
    for (int e = threadIdx.x; e < kGM * kGK; e += 256) {
      const int mm = a_k_fast ? e / kGK : e % kGM, kk = a_k_fast ? e % kGK : e / kGM;
      const int m = m0 + mm, k = k0 + kk;
      sA[kk][mm] = (m < M && k < kend) ? A[m * sa_m + k * sa_k] : 0.f;
    }
    for (int e = threadIdx.x; e < kGN * kGK; e += 256) {
      const int nn = b_n_fast ? e % kGN : e / kGK, kk = b_n_fast ? e / kGN : e % kGK;
      const int n = n0 + nn, k = k0 + kk;
      sB[kk][nn] = (n < N && k < kend) ? B[k * sb_k + n * sb_n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (partial)
        partial[(size_t)blockIdx.z * M * N + (size_t)m * N + n] = acc[i][j];
      else
        C[(size_t)m * ldc + n] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N,
                                                            float* __restrict__ C, long long ldc) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)M * N) return;
  double acc = 0.0;
  for (int s = 0; s < splits; ++s) acc += (double)partial[(size_t)s * M * N + e];
  C[(size_t)(e / N) * ldc + (e % N)] = (float)acc;
}

}  // namespace cl3d

using namespace cl3d;

extern "C" size_t cl3d_sgemm_workspace_bytes(int M, int N, int splitk) {
  return splitk > 1 ? sizeof(float) * (size_t)splitk * M * N : 0;
}

extern "C" int cl3d_sgemm(const float* a, long long sa_m, long long sa_k, const float* b, long long sb_k,
                          long long sb_n, int M, int N, int K, float* c, long long ldc, int splitk, void* workspace,
                          size_t workspace_bytes, cl3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CL3D_REQUIRE(a && b && c && M >= 1 && N >= 1 && K >= 1 && ldc >= N, "cl3d_sgemm: bad arguments");
  if (splitk < 1) splitk = 1;
  if (splitk > K) splitk = K;
  CL3D_REQUIRE(splitk <= 65535, "cl3d_sgemm: splitk too large");
  int kps = ceil_div(K, splitk);
  kps = ceil_div(kps, kGK) * kGK;
  splitk = ceil_div(K, kps);
  float* partial = nullptr;
  if (splitk > 1) {
    if (!workspace || workspace_bytes < cl3d_sgemm_workspace_bytes(M, N, splitk)) {
      set_error("cl3d_sgemm: split-K workspace too small");
      return CL3D_ERR_WORKSPACE;
    }
    partial = (float*)workspace;
  }
  dim3 grid(ceil_div(N, kGN), ceil_div(M, kGM), splitk);
  sgemm_strided_kernel<<<grid, 256, 0, stream>>>(a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial); CL3D_LAUNCHED(1);
  if (splitk > 1) {
    const long long total = (long long)M * N;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(partial, splitk, M, N, c, ldc); CL3D_LAUNCHED(1);
  }
  return check_launch("sgemm_strided_kernel");
}
