// pool.cu -- fused neighbourhood max-pool (the body of the reference's MaskedMaxPool,
//   /root/reference/pytorch/ops/pt_custom_ops/pt_utils.py:188-202: group_points -> F.max_pool2d over K)
// and its gradient, without the (B,C,M,K) tensor.  Also used by MaskedUpsample(mode='max').
//
// forward : warp per query, lane = channel, neighbour rows of the point-major feature matrix streamed with
//           coalesced loads; max over ALL K slots (the reference ignores masks here; cyclic padding duplicates
//           cannot change a max), first arg-max slot kept (max_pool2d's tie rule) for the backward.
// backward: grad_f[b, c, idx[b,q,arg]] += grad_out[b,c,q]  -- one fp32 red.add row per query into a zeroed
//           point-major buffer (the reference scatters K times as many atomics through group_points_grad).
#include "common.cuh"

namespace cl3d {

constexpr int kPoolWarps = 8;
constexpr int kPoolTile = 32;
constexpr int kPoolU = 4;

template <int CI>
__global__ void __launch_bounds__(kPoolWarps * 32) gather_max_fwd_kernel(const float* __restrict__ feat_pm,
                                                                         const int* __restrict__ idx, int N, int M,
                                                                         int K, int C, int Cp, float* __restrict__ out,
                                                                         unsigned char* __restrict__ arg) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * 32 * CI;
  const int chunkC = min(32 * CI, Cp - c0);
  unsigned* s_off = reinterpret_cast<unsigned*>(smem) + (size_t)warp * K;
  float* s_out = reinterpret_cast<float*>(smem + align_up((size_t)kPoolWarps * K * 4, 16));
  const int tiles_per_cloud = (M + kPoolTile - 1) / kPoolTile;
  const int b = blockIdx.x / tiles_per_cloud;
  const int q0 = (blockIdx.x % tiles_per_cloud) * kPoolTile;
  const float* base = feat_pm + (size_t)b * N * Cp + c0 + lane;
  for (int ql = warp; ql < kPoolTile; ql += kPoolWarps) {
    const int q = q0 + ql;
    float m[CI];
    int km[CI];
#pragma unroll
    for (int i = 0; i < CI; ++i) { m[i] = -INFINITY; km[i] = 0; }
    if (q < M) {
      const size_t gq = (size_t)b * M + q;
      for (int k = lane; k < K; k += 32) s_off[k] = (unsigned)idx[gq * K + k] * (unsigned)Cp;
      __syncwarp();
      int k0 = 0;
      for (; k0 + kPoolU <= K; k0 += kPoolU) {
        float v[kPoolU][CI];
#pragma unroll
        for (int u = 0; u < kPoolU; ++u) {
          const float* row = base + s_off[k0 + u];
#pragma unroll
          for (int i = 0; i < CI; ++i) v[u][i] = __ldg(row + 32 * i);
        }
#pragma unroll
        for (int u = 0; u < kPoolU; ++u)
#pragma unroll
          for (int i = 0; i < CI; ++i) {
            const bool gt = v[u][i] > m[i];
            m[i] = gt ? v[u][i] : m[i];
            km[i] = gt ? k0 + u : km[i];
          }
      }
      for (; k0 < K; ++k0) {
        const float* row = base + s_off[k0];
#pragma unroll
        for (int i = 0; i < CI; ++i) {
          const float t = __ldg(row + 32 * i);
          const bool gt = t > m[i];
          m[i] = gt ? t : m[i];
          km[i] = gt ? k0 : km[i];
        }
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < CI; ++i)
        if (lane + 32 * i < chunkC) arg[gq * Cp + c0 + lane + 32 * i] = (unsigned char)km[i];
    }
#pragma unroll
    for (int i = 0; i < CI; ++i) s_out[(size_t)(lane + 32 * i) * (kPoolTile + 1) + ql] = m[i];
  }
  __syncthreads();
  const int q = q0 + lane;
  for (int cl = warp; cl < 32 * CI; cl += kPoolWarps) {
    const int c = c0 + cl;
    if (c >= C) break;
    if (q < M) out[((size_t)b * C + c) * M + q] = s_out[(size_t)cl * (kPoolTile + 1) + lane];
  }
}

// grad_pm[b, idx[b,q,arg[b,q,c]], c] += grad_out[b,c,q]
__global__ void __launch_bounds__(256) gather_max_bwd_kernel(const float* __restrict__ grad_out,
                                                             const int* __restrict__ idx,
                                                             const unsigned char* __restrict__ arg, int N, int M, int K,
                                                             int C, int Cp, float* __restrict__ grad_pm) {
  extern __shared__ float s_g[];  // [32 queries][C + 1]
  const int tiles_per_cloud = (M + 31) / 32;
  const int b = blockIdx.x / tiles_per_cloud;
  const int q0 = (blockIdx.x % tiles_per_cloud) * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < C; c += 8) {
    const int q = q0 + lane;
    s_g[(size_t)lane * (C + 1) + c] = q < M ? grad_out[((size_t)b * C + c) * M + q] : 0.f;
  }
  __syncthreads();
  for (int ql = warp; ql < 32; ql += 8) {
    const int q = q0 + ql;
    if (q >= M) continue;
    const size_t gq = (size_t)b * M + q;
    for (int c = lane; c < C; c += 32) {
      const int j = idx[gq * K + arg[gq * Cp + c]];
      atomicAdd(grad_pm + ((size_t)b * N + j) * Cp + c, s_g[(size_t)ql * (C + 1) + c]);
    }
  }
}

template <int CI>
static int launch_pool_fwd(const float* feat_pm, const int* idx, int B, int N, int M, int K, int C, int Cp, float* out,
                           unsigned char* arg, cudaStream_t stream) {
  const size_t smem = align_up((size_t)kPoolWarps * K * 4, 16) + (size_t)32 * CI * (kPoolTile + 1) * 4;
  static std::atomic<unsigned long long> seen{0};
  allow_big_smem(gather_max_fwd_kernel<CI>, seen);
  dim3 grid(B * ceil_div(M, kPoolTile), ceil_div(Cp, 32 * CI));
  gather_max_fwd_kernel<CI><<<grid, kPoolWarps * 32, smem, stream>>>(feat_pm, idx, N, M, K, C, Cp, out, arg);
  CL3D_LAUNCHED(1);
  return check_launch("gather_max_fwd_kernel");
}

}  // namespace cl3d

using namespace cl3d;

extern "C" int cl3d_gather_max_fwd(const float* feat_pm, const int* idx, int B, int N, int M, int K, int C, float* out,
                                   unsigned char* arg, cl3d_stream_t stream_) {
  CL3D_REQUIRE(feat_pm && idx && out && arg && B >= 0 && N >= 1 && M >= 1 && K >= 1 && K <= 255 && C >= 1,
               "cl3d_gather_max_fwd: bad arguments (K <= 255)");
  if (B == 0) return CL3D_OK;
  const int Cp = padded_channels(C);
  int ci = ceil_div(Cp, 32);
  if (ci > 4) ci = 4;
  cudaStream_t s = (cudaStream_t)stream_;
  switch (ci) {
    case 1: return launch_pool_fwd<1>(feat_pm, idx, B, N, M, K, C, Cp, out, arg, s);
    case 2: return launch_pool_fwd<2>(feat_pm, idx, B, N, M, K, C, Cp, out, arg, s);
    case 3: return launch_pool_fwd<3>(feat_pm, idx, B, N, M, K, C, Cp, out, arg, s);
    default: return launch_pool_fwd<4>(feat_pm, idx, B, N, M, K, C, Cp, out, arg, s);
  }
}

extern "C" int cl3d_gather_max_bwd(const float* grad_out, const int* idx, const unsigned char* arg, int B, int N,
                                   int M, int K, int C, float* grad_pm, cl3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CL3D_REQUIRE(grad_out && idx && arg && grad_pm && B >= 0 && N >= 1 && M >= 1 && K >= 1 && C >= 1,
               "cl3d_gather_max_bwd: bad arguments");
  if (B == 0) return CL3D_OK;
  const int Cp = padded_channels(C);
  cudaMemsetAsync(grad_pm, 0, sizeof(float) * (size_t)B * N * Cp, stream);
  const size_t smem = (size_t)32 * (C + 1) * sizeof(float);
  static std::atomic<unsigned long long> seen{0};
  allow_big_smem(gather_max_bwd_kernel, seen);
  gather_max_bwd_kernel<<<B * ceil_div(M, 32), 256, smem, stream>>>(grad_out, idx, arg, N, M, K, C, Cp, grad_pm);
  CL3D_LAUNCHED(1);
  return check_launch("gather_max_bwd_kernel");
}
