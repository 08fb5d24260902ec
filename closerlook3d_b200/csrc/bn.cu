// bn.cu -- the out_transform of PosPool / AdaptiveWeight / PseudoGrid: BatchNorm1d + ReLU on the
//          channel-major aggregated tensor (B,C,M), forward and backward (sm_100a).
//
// Replaces nn.Sequential(nn.BatchNorm1d(C, momentum), nn.ReLU(inplace=True)) of
//   /root/reference/pytorch/models/local_aggregation_operators.py:43-45,110 (and :166-168, :363-365).
// Batch statistics come from the per-tile partial sums the aggregation kernel already produced, so the
// forward costs one read + one write of (B,C,M); the backward is two passes (statistics, apply) and writes
// d(loss)/d(agg) directly in the point-major layout the gather-form aggregation backward consumes.
#include "common.cuh"

namespace cl3d {

constexpr int kBnTile = 32;

// mean / invstd from the partial sums; running-stat update as nn.BatchNorm1d (unbiased variance).
// One CTA of 1024 threads handles 32 channels (32 tile-lanes each), deterministic.
// One CTA = 8 channels (one 32-byte sector of a partial row) x 128 tile slices: C/8 CTAs, so the whole partial
// array is in flight at once (the kernel sits on the forward's critical path and is pure load latency).
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partial, int ntiles, int C,
                                                           double count, float eps, float momentum, int training,
                                                           float* __restrict__ running_mean,
                                                           float* __restrict__ running_var,
                                                           float* __restrict__ save_stats) {
  __shared__ double s1[128][9], s2[128][9];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + tx;
  double a1 = 0.0, a2 = 0.0;
  if (training && c < C) {
    double b1 = 0.0, b2 = 0.0, c1 = 0.0, c2 = 0.0, d1 = 0.0, d2 = 0.0;
    int t = ty;
    for (; t + 384 < ntiles; t += 512) {  // four independent load streams per thread
      a1 += (double)partial[((size_t)t * 2 + 0) * C + c];
      a2 += (double)partial[((size_t)t * 2 + 1) * C + c];
      b1 += (double)partial[((size_t)(t + 128) * 2 + 0) * C + c];
      b2 += (double)partial[((size_t)(t + 128) * 2 + 1) * C + c];
      c1 += (double)partial[((size_t)(t + 256) * 2 + 0) * C + c];
      c2 += (double)partial[((size_t)(t + 256) * 2 + 1) * C + c];
      d1 += (double)partial[((size_t)(t + 384) * 2 + 0) * C + c];
      d2 += (double)partial[((size_t)(t + 384) * 2 + 1) * C + c];
    }
    for (; t < ntiles; t += 128) {
      a1 += (double)partial[((size_t)t * 2 + 0) * C + c];
      a2 += (double)partial[((size_t)t * 2 + 1) * C + c];
    }
    a1 = (a1 + b1) + (c1 + d1);
    a2 = (a2 + b2) + (c2 + d2);
  }
  s1[ty][tx] = a1;
  s2[ty][tx] = a2;
  __syncthreads();
  // fixed-order tree over the 128 slices (deterministic)
  for (int h = 64; h >= 1; h >>= 1) {
    if (ty < h) {
      s1[ty][tx] += s1[ty + h][tx];
      s2[ty][tx] += s2[ty + h][tx];
    }
    __syncthreads();
  }
  if (ty == 0 && c < C) {
    if (training) {
      const double sum = s1[0][tx], sq = s2[0][tx];
      const double mean = sum / count;
      double var = sq / count - mean * mean;
      if (var < 0.0) var = 0.0;
      save_stats[c] = (float)mean;
      save_stats[C + c] = (float)(1.0 / sqrt(var + (double)eps));
      if (running_mean && running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
        running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unbiased);
      }
    } else {
      save_stats[c] = running_mean[c];
      save_stats[C + c] = 1.0f / sqrtf(running_var[c] + eps);
    }
  }
}

// y = relu((x - mean) * invstd * gamma + beta), channel-major, float4 along M when aligned
__global__ void __launch_bounds__(256) bn_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, int C, int M,
                                                          float* __restrict__ y) {
  const int row = blockIdx.x;  // b*C + c
  const int c = row % C;
  const float mean = stats[c], invstd = stats[C + c];
  const float sc = invstd * gamma[c], sh = beta[c];
  const float* xr = x + (size_t)row * M;
  float* yr = y + (size_t)row * M;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < M; i += gridDim.y * blockDim.x) {
    const float v = __fmaf_rn(__fsub_rn(xr[i], mean), sc, sh);  // the backward recomputes exactly this
    yr[i] = v > 0.f ? v : 0.f;
  }
}

// Backward, tiles of 128 consecutive queries of one cloud: a warp owns a channel row segment (512 bytes, one
// float4 per lane when M % 4 == 0), so both passes stream (B,C,M) with full-width, page-friendly accesses.
constexpr int kBwdTile = 128;
constexpr int kBwdCh = 32;  // channels staged per shared-memory transpose round in the apply pass

__device__ __forceinline__ void load4(const float* __restrict__ p, int q, int M, bool vec, float (&v)[4]) {
  if (vec && q + 3 < M) {
    const float4 t = *reinterpret_cast<const float4*>(p + q);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t] = (q + t < M) ? p[q + t] : 0.f;
  }
}

// pass 1: per tile and channel: sum(gy), sum(gy * xhat)
__global__ void __launch_bounds__(256) bn_relu_bwd_stats_kernel(const float* __restrict__ grad_y,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ stats,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, int C, int M,
                                                                float* __restrict__ partial) {
  const int tiles_per_cloud = (M + kBwdTile - 1) / kBwdTile;
  const int b = blockIdx.x / tiles_per_cloud;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = (blockIdx.x % tiles_per_cloud) * kBwdTile + lane * 4;
  const bool vec = (M & 3) == 0;
  for (int c = warp; c < C; c += 8) {
    const size_t row = ((size_t)b * C + c) * M;
    float xv[4], gv[4];
    load4(x + row, q, M, vec, xv);
    load4(grad_y + row, q, M, vec, gv);
    const float mean = stats[c], invstd = stats[C + c], sc = invstd * gamma[c], sh = beta[c];
    float gy = 0.f, gx = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (q + t < M) {
        const float xc = __fsub_rn(xv[t], mean);
        const float yv = __fmaf_rn(xc, sc, sh);  // bit-identical to the forward's y
        const float g = yv > 0.f ? gv[t] : 0.f;
        gy += g;
        gx = fmaf(g, xc * invstd, gx);
      }
    }
    const float s1 = warp_sum(gy), s2 = warp_sum(gx);
    if (lane == 0) {
      partial[((size_t)blockIdx.x * 2 + 0) * C + c] = s1;
      partial[((size_t)blockIdx.x * 2 + 1) * C + c] = s2;
    }
  }
}

// pass 2: g = gamma*invstd*(gy - mean(gy) - xhat*mean(gy*xhat)) (training) or gamma*invstd*gy (eval), written
// point-major (B,M,Cp) through a shared-memory transpose, kBwdCh channels per round; padding channels = 0.
__global__ void __launch_bounds__(256) bn_relu_bwd_apply_kernel(const float* __restrict__ grad_y,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ stats,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta,
                                                                const float* __restrict__ dgb, int C, int Cp, int M,
                                                                float inv_count, int training,
                                                                float* __restrict__ g_pm) {
  __shared__ float s_tile[kBwdTile][kBwdCh + 1];
  const int tiles_per_cloud = (M + kBwdTile - 1) / kBwdTile;
  const int b = blockIdx.x / tiles_per_cloud;
  const int q0 = (blockIdx.x % tiles_per_cloud) * kBwdTile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = q0 + lane * 4;
  const bool vec = (M & 3) == 0;
  const int nq = min(kBwdTile, M - q0);
  for (int cb = 0; cb < Cp; cb += kBwdCh) {
    for (int cl = warp; cl < kBwdCh; cl += 8) {
      const int c = cb + cl;
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      if (c < C) {
        const size_t row = ((size_t)b * C + c) * M;
        float xv[4], gv[4];
        load4(x + row, q, M, vec, xv);
        load4(grad_y + row, q, M, vec, gv);
        const float mean = stats[c], invstd = stats[C + c], ga = gamma[c], sc = invstd * ga, sh = beta[c];
        const float mg = training ? dgb[C + c] * inv_count : 0.f, mgx = training ? dgb[c] * inv_count : 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float xc = __fsub_rn(xv[t], mean);
          const float yv = __fmaf_rn(xc, sc, sh);
          const float gy = yv > 0.f ? gv[t] : 0.f;
          g[t] = sc * (gy - mg - xc * invstd * mgx);
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) s_tile[lane * 4 + t][cl] = g[t];
    }
    __syncthreads();
    const int ncb = min(kBwdCh, Cp - cb);
    for (int e = threadIdx.x; e < nq * ncb; e += blockDim.x) {
      const int ql = e / ncb, cl = e % ncb;
      g_pm[((size_t)b * M + q0 + ql) * Cp + cb + cl] = s_tile[ql][cl];
    }
    __syncthreads();
  }
}

// dgamma[c] = sum gy*xhat, dbeta[c] = sum gy from the partials: reuse the generic fixed-order reduction
__global__ void __launch_bounds__(1024) bn_reduce2_kernel(const float* __restrict__ partial, int ntiles, int C,
                                                          float* __restrict__ dgb /*(2,C): dgamma, dbeta*/) {
  __shared__ double s1[32][33], s2[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double a1 = 0.0, a2 = 0.0;
  if (c < C) {
    double b1 = 0.0, b2 = 0.0, c1 = 0.0, c2 = 0.0, d1 = 0.0, d2 = 0.0;
    int t = ty;
    for (; t + 96 < ntiles; t += 128) {
      a1 += (double)partial[((size_t)t * 2 + 0) * C + c];
      a2 += (double)partial[((size_t)t * 2 + 1) * C + c];
      b1 += (double)partial[((size_t)(t + 32) * 2 + 0) * C + c];
      b2 += (double)partial[((size_t)(t + 32) * 2 + 1) * C + c];
      c1 += (double)partial[((size_t)(t + 64) * 2 + 0) * C + c];
      c2 += (double)partial[((size_t)(t + 64) * 2 + 1) * C + c];
      d1 += (double)partial[((size_t)(t + 96) * 2 + 0) * C + c];
      d2 += (double)partial[((size_t)(t + 96) * 2 + 1) * C + c];
    }
    for (; t < ntiles; t += 32) {
      a1 += (double)partial[((size_t)t * 2 + 0) * C + c];
      a2 += (double)partial[((size_t)t * 2 + 1) * C + c];
    }
    a1 = (a1 + b1) + (c1 + d1);
    a2 = (a2 + b2) + (c2 + d2);
  }
  s1[ty][tx] = a1;
  s2[ty][tx] = a2;
  __syncthreads();
  if (ty == 0 && c < C) {
    double sgy = 0.0, sgx = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      sgy += s1[i][tx];
      sgx += s2[i][tx];
    }
    dgb[c] = (float)sgx;      // dgamma
    dgb[C + c] = (float)sgy;  // dbeta
  }
}

}  // namespace cl3d

using namespace cl3d;

extern "C" int cl3d_bn_finalize(const float* bn_partial, int ntiles, int C, long long count, float eps, float momentum,
                                int training, float* running_mean, float* running_var, float* save_stats,
                                cl3d_stream_t stream_) {
  CL3D_REQUIRE(C >= 1 && save_stats, "cl3d_bn_finalize: bad arguments");
  CL3D_REQUIRE(!training || (bn_partial && ntiles >= 1 && count >= 1), "cl3d_bn_finalize: training needs partial sums");
  CL3D_REQUIRE(training || (running_mean && running_var), "cl3d_bn_finalize: eval needs running statistics");
  bn_finalize_kernel<<<ceil_div(C, 8), 1024, 0, (cudaStream_t)stream_>>>(bn_partial, ntiles, C, (double)count, eps,
                                                                         momentum, training, running_mean,
                                                                         running_var, save_stats); CL3D_LAUNCHED(1);
  return check_launch("bn_finalize_kernel");
}

extern "C" int cl3d_bn_relu_fwd(const float* x, const float* save_stats, const float* gamma, const float* beta, int B,
                                int C, int M, float* y, cl3d_stream_t stream_) {
  CL3D_REQUIRE(x && save_stats && gamma && beta && y && B >= 0 && C >= 1 && M >= 1, "cl3d_bn_relu_fwd: bad arguments");
  if (B == 0) return CL3D_OK;
  int chunks = ceil_div(M, 1024);
  chunks = chunks < 1 ? 1 : (chunks > 65535 ? 65535 : chunks);
  dim3 grid(B * C, chunks);
  bn_relu_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(x, save_stats, gamma, beta, C, M, y); CL3D_LAUNCHED(1);
  return check_launch("bn_relu_fwd_kernel");
}

extern "C" int cl3d_bn_relu_bwd(const float* grad_y, const float* x, const float* save_stats, const float* gamma,
                                const float* beta, int B, int C, int M, int training, float* partial,
                                float* dgamma_dbeta, float* g_pm, cl3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CL3D_REQUIRE(grad_y && x && save_stats && gamma && beta && partial && dgamma_dbeta && g_pm,
               "cl3d_bn_relu_bwd: null pointer");
  CL3D_REQUIRE(B >= 0 && C >= 1 && M >= 1, "cl3d_bn_relu_bwd: bad sizes");
  if (B == 0) return CL3D_OK;
  const int ntiles = B * ceil_div(M, kBwdTile);  // <= cl3d_agg_num_tiles(B, M): the caller's partial buffer fits
  const int Cp = padded_channels(C);
  bn_relu_bwd_stats_kernel<<<ntiles, 256, 0, stream>>>(grad_y, x, save_stats, gamma, beta, C, M, partial); CL3D_LAUNCHED(1);
  bn_reduce2_kernel<<<ceil_div(C, 32), 1024, 0, stream>>>(partial, ntiles, C, dgamma_dbeta); CL3D_LAUNCHED(1);
  bn_relu_bwd_apply_kernel<<<ntiles, 256, 0, stream>>>(grad_y, x, save_stats, gamma, beta, dgamma_dbeta, C, Cp, M,
                                                       1.0f / (float)((long long)B * M), training, g_pm); CL3D_LAUNCHED(1);
  return check_launch("bn_relu_bwd kernels");
}
