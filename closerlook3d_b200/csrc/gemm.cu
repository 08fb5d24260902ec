// gemm.cu -- fp32 FMA GEMM with arbitrary element strides and optional split-K (sm_100a).
//
// Used for the per-point products of the fused PointWiseMLP (see pwmlp.cu): the reference's per-neighbour
// 1x1 conv over [dp; f_i; f_j - f_i] (/root/reference/pytorch/models/local_aggregation_operators.py:254-257,
// 288-295) becomes ONE per-point product (B*N x (C+3)) x ((C+3) x 2*Cout), plus its two gradients.  These are
// skinny problems (N, K ~ 75..150): far too small to fill a tcgen05 tile, and TF32/BF16 operands would break
// the fp32 1e-5 parity bar, so they stay vectorised fp32 FMA (BASELINE.json north_star).
//
//   C[m][n] = sum_k A[m*sa_m + k*sa_k] * B[k*sb_k + n*sb_n]
// CTA tile 128 x (16*TN), k tile 16, 256 threads, 8 x TN outputs per thread (TN chosen from N so that skinny
// outputs waste few columns), operands transposed into shared memory as [k][m] / [k][n] (LDS.128 for the A
// fragment), next k tile prefetched into registers while the current one is multiplied.  With splitk > 1 the k
// range is divided among gridDim.z CTAs that write partial tiles, reduced afterwards in a fixed order.
#include "common.cuh"

namespace cl3d {

constexpr int kBM = 128, kBK = 16;

template <int TN, bool VEC>
__global__ void __launch_bounds__(256) sgemm_tiled_kernel(const float* __restrict__ A, long long sa_m, long long sa_k,
                                                          const float* __restrict__ B, long long sb_k,
                                                          long long sb_n, int M, int N, int K, int k_per_split,
                                                          float* __restrict__ C, long long ldc,
                                                          float* __restrict__ partial) {
  constexpr int BN = 16 * TN;
  __shared__ __align__(16) float sA[kBK][kBM + 4];
  __shared__ __align__(16) float sB[kBK][BN + 4];
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool a_k_fast = (sa_k == 1), b_k_fast = (sb_k == 1);
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  // VEC: every operand is read as float4 along its unit-stride dimension (host guarantees 16-byte alignment,
  // strides and extents that are multiples of 4, so a vector is entirely inside or entirely outside)
  constexpr int NVB = (BN * kBK / 4 + 255) / 256;  // float4 of the B tile per thread
  float ra[8], rb[VEC ? NVB * 4 : TN];

  auto load_tiles = [&](int k0) {
    if constexpr (VEC) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int e = threadIdx.x + 256 * r;  // 512 float4 of the A tile
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_k_fast) {
          const int mm = e >> 2, kk = (e & 3) * 4;
          const int m = m0 + mm, k = k0 + kk;
          if (m < M && k < kend) v = __ldg(reinterpret_cast<const float4*>(A + m * sa_m + k));
        } else {
          const int kk = e >> 5, mm = (e & 31) * 4;
          const int m = m0 + mm, k = k0 + kk;
          if (m < M && k < kend) v = __ldg(reinterpret_cast<const float4*>(A + k * sa_k + m));
        }
        ra[r * 4 + 0] = v.x; ra[r * 4 + 1] = v.y; ra[r * 4 + 2] = v.z; ra[r * 4 + 3] = v.w;
      }
#pragma unroll
      for (int r = 0; r < NVB; ++r) {
        const int e = threadIdx.x + 256 * r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < BN * kBK / 4) {
          if (b_k_fast) {
            const int nn = e >> 2, kk = (e & 3) * 4;
            const int n = n0 + nn, k = k0 + kk;
            if (n < N && k < kend) v = __ldg(reinterpret_cast<const float4*>(B + n * sb_n + k));
          } else {
            const int kk = e / (BN / 4), nn = (e % (BN / 4)) * 4;
            const int n = n0 + nn, k = k0 + kk;
            if (n < N && k < kend) v = __ldg(reinterpret_cast<const float4*>(B + k * sb_k + n));
          }
        }
        rb[r * 4 + 0] = v.x; rb[r * 4 + 1] = v.y; rb[r * 4 + 2] = v.z; rb[r * 4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int e = threadIdx.x + 256 * r;
        const int mm = a_k_fast ? e / kBK : e % kBM, kk = a_k_fast ? e % kBK : e / kBM;
        const int m = m0 + mm, k = k0 + kk;
        ra[r] = (m < M && k < kend) ? __ldg(A + m * sa_m + k * sa_k) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < TN; ++r) {
        const int e = threadIdx.x + 256 * r;
        const int nn = b_k_fast ? e / kBK : e % BN, kk = b_k_fast ? e % kBK : e / BN;
        const int n = n0 + nn, k = k0 + kk;
        rb[r] = (n < N && k < kend) ? __ldg(B + k * sb_k + n * sb_n) : 0.f;
      }
    }
  };
  auto store_tiles = [&]() {
    if constexpr (VEC) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int e = threadIdx.x + 256 * r;
        if (a_k_fast) {
          const int mm = e >> 2, kk = (e & 3) * 4;
#pragma unroll
          for (int t = 0; t < 4; ++t) sA[kk + t][mm] = ra[r * 4 + t];
        } else {
          const int kk = e >> 5, mm = (e & 31) * 4;
          *reinterpret_cast<float4*>(&sA[kk][mm]) = make_float4(ra[r * 4], ra[r * 4 + 1], ra[r * 4 + 2], ra[r * 4 + 3]);
        }
      }
#pragma unroll
      for (int r = 0; r < NVB; ++r) {
        const int e = threadIdx.x + 256 * r;
        if (e < BN * kBK / 4) {
          if (b_k_fast) {
            const int nn = e >> 2, kk = (e & 3) * 4;
#pragma unroll
            for (int t = 0; t < 4; ++t) sB[kk + t][nn] = rb[r * 4 + t];
          } else {
            const int kk = e / (BN / 4), nn = (e % (BN / 4)) * 4;
            *reinterpret_cast<float4*>(&sB[kk][nn]) = make_float4(rb[r * 4], rb[r * 4 + 1], rb[r * 4 + 2], rb[r * 4 + 3]);
          }
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int e = threadIdx.x + 256 * r;
        const int mm = a_k_fast ? e / kBK : e % kBM, kk = a_k_fast ? e % kBK : e / kBM;
        sA[kk][mm] = ra[r];
      }
#pragma unroll
      for (int r = 0; r < TN; ++r) {
        const int e = threadIdx.x + 256 * r;
        const int nn = b_k_fast ? e / kBK : e % BN, kk = b_k_fast ? e % kBK : e / BN;
        sB[kk][nn] = rb[r];
      }
    }
  };

  load_tiles(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += kBK) {
    store_tiles();
    __syncthreads();
    if (k0 + kBK < kend) load_tiles(k0 + kBK);  // prefetch the next k tile into registers
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[kk][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = sB[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
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
  const size_t stride = (size_t)M * N;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int s = 0;
  for (; s + 4 <= splits; s += 4) {
    a0 += (double)partial[(size_t)(s + 0) * stride + e];
    a1 += (double)partial[(size_t)(s + 1) * stride + e];
    a2 += (double)partial[(size_t)(s + 2) * stride + e];
    a3 += (double)partial[(size_t)(s + 3) * stride + e];
  }
  for (; s < splits; ++s) a0 += (double)partial[(size_t)s * stride + e];
  C[(size_t)(e / N) * ldc + (e % N)] = (float)((a0 + a1) + (a2 + a3));
}

template <int TN>
static void launch_sgemm(dim3 grid, cudaStream_t stream, const float* a, long long sa_m, long long sa_k, const float* b,
                         long long sb_k, long long sb_n, int M, int N, int K, int kps, float* c, long long ldc,
                         float* partial) {
  // float4 path: unit stride along one dimension of each operand, everything else a multiple of 4 floats
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool va = al16(a) && ((sa_k == 1 && sa_m % 4 == 0 && K % 4 == 0 && kps % 4 == 0) ||
                              (sa_m == 1 && sa_k % 4 == 0 && M % 4 == 0));
  const bool vb = al16(b) && ((sb_k == 1 && sb_n % 4 == 0 && K % 4 == 0 && kps % 4 == 0) ||
                              (sb_n == 1 && sb_k % 4 == 0 && N % 4 == 0));
  if (va && vb && (sa_k == 1 || sa_m == 1) && (sb_k == 1 || sb_n == 1))
    sgemm_tiled_kernel<TN, true><<<grid, 256, 0, stream>>>(a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial);
  else
    sgemm_tiled_kernel<TN, false><<<grid, 256, 0, stream>>>(a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial);
  CL3D_LAUNCHED(1);
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
  kps = ceil_div(kps, kBK) * kBK;
  splitk = ceil_div(K, kps);
  float* partial = nullptr;
  if (splitk > 1) {
    if (!workspace || workspace_bytes < cl3d_sgemm_workspace_bytes(M, N, splitk)) {
      set_error("cl3d_sgemm: split-K workspace too small");
      return CL3D_ERR_WORKSPACE;
    }
    partial = (float*)workspace;
  }
  // columns per thread: spread N over the fewest 144-wide column tiles, then round each tile up to 16*TN
  const int nct = ceil_div(N, 144);
  int tn = ceil_div(ceil_div(N, nct), 16);
  if (tn < 1) tn = 1;
  if (tn > 9) tn = 9;
  dim3 grid(ceil_div(N, 16 * tn), ceil_div(M, kBM), splitk);
  switch (tn) {
    case 1: launch_sgemm<1>(grid, stream, a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial); break;
    case 2: launch_sgemm<2>(grid, stream, a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial); break;
    case 3: launch_sgemm<3>(grid, stream, a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial); break;
    case 4: launch_sgemm<4>(grid, stream, a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial); break;
    case 5: launch_sgemm<5>(grid, stream, a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial); break;
    case 6: launch_sgemm<6>(grid, stream, a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial); break;
    case 7: launch_sgemm<7>(grid, stream, a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial); break;
    case 8: launch_sgemm<8>(grid, stream, a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial); break;
    default: launch_sgemm<9>(grid, stream, a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, kps, c, ldc, partial); break;
  }
  if (splitk > 1) {
    const long long total = (long long)M * N;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(partial, splitk, M, N, c, ldc);
    CL3D_LAUNCHED(1);
  }
  return check_launch("sgemm_tiled_kernel");
}
