// pwmlp.cu -- fused Point-wise MLP local aggregation (feature_type 'dp_fi_df', num_mlps 1, reduction max).
//
// Reference: /root/reference/pytorch/models/local_aggregation_operators.py:254-257 (conv + BatchNorm2d + ReLU),
// :288-303 (input assembly, max over K).  With one conv layer
//     y[o,q,k] = Wp[o]·dp_k + Wc[o]·f_i + Wr[o]·(f_{j_k} - f_i),   f_i = f of slot 0 (nearest neighbour)
//              = Wp[o]·dp_k + A[j_0][o] + Bv[j_k][o],   A = f (Wc-Wr)^T,  Bv = f Wr^T   (per POINT, gemm.cu)
// so the per-neighbour GEMM over (B,147,M,K) becomes two per-point products plus a gather-add; BatchNorm2d
// statistics over all B*M*K positions (padding slots included, as in the reference) are accumulated on the
// fly, and the max over K commutes with the monotone BN+ReLU:  out = relu(sc*(sc>=0 ? max_k y : min_k y)+sh).
//
// forward : pwmlp_fwd_kernel  (warp per query, TMA bulk-copy gather of Bv rows, writes max/min/argmax/argmin
//                              + per-tile BN partial sums)  ->  bn_finalize  ->  pwmlp_out_kernel
// backward: pwmlp_bwd_stats_kernel (sum dz, sum dz*yhat)  ->  pwmlp_bwd_kernel (query-major gather again,
//           recomputes y, emits d/dA, d/dBv (fp32 red.add into the point-major buffer), d/dWp partials)
#include "common.cuh"

namespace cl3d {

constexpr int kPWWarps = 8;
constexpr int kPWTile = 32;
constexpr int kPWStageBytes = 8192;
constexpr int kPWMaxCI = 4;  // 128 output channels per CTA chunk

struct PwArgs {
  const float* ab_pm;        // (B,N,2*Cop): row = [A (Cop) | Bv (Cop)]
  const float* wp;           // (Cout,3)
  const float* query_xyz;    // (B,M,3)
  const float* support_xyz;  // (B,N,3)
  const int* idx;            // (B,M,K)
  float* ymax;               // (B,Cout,M)
  float* ymin;               // (B,Cout,M)
  unsigned short* arg;       // (B,M,Cop): argmax | argmin << 8
  float* partial;            // fwd: (ntiles,2,Cout) ; bwd: (ntiles,3,Cout) dWp partials
  // backward only
  const float* grad_out;     // (B,Cout,M)
  const float* out;          // (B,Cout,M)
  const float* stats;        // (2,Cout) mean, invstd
  const float* gamma;        // (Cout)
  const float* dgb;          // (2,Cout): sum dz*yhat, sum dz
  float* grad_ab_pm;         // (B,N,2*Cop), zero-initialised by the caller
  int B, N, M, K, Cout, Cop;
  float inv_radius, inv_count;
  int rows_per_stage, ntiles;
};

struct PwSmem {
  size_t stage_off, a_off, dp_off, idx_off, t0_off, t1_off, arg_off, red_off, bar_off, total;
};
__host__ __device__ inline PwSmem pw_smem(int K, int chunk) {
  PwSmem L;
  size_t o = 0;
  L.stage_off = o; o += (size_t)kPWWarps * kPWStageBytes;
  L.a_off = o;     o += (size_t)kPWWarps * chunk * sizeof(float);
  L.dp_off = o;    o += (size_t)kPWWarps * K * sizeof(float4);
  L.idx_off = o;   o += (size_t)kPWWarps * K * sizeof(int);
  o = align_up(o, 16);
  L.t0_off = o;    o += (size_t)chunk * (kPWTile + 1) * sizeof(float);
  L.t1_off = o;    o += (size_t)chunk * (kPWTile + 1) * sizeof(float);
  L.arg_off = o;   o += (size_t)kPWTile * chunk * sizeof(unsigned short);
  o = align_up(o, 16);
  L.red_off = o;   o += (size_t)kPWWarps * 3 * chunk * sizeof(float);
  o = align_up(o, 16);
  L.bar_off = o;   o += (size_t)kPWWarps * sizeof(uint64_t);
  L.total = o;
  return L;
}

// =================================================================================================
// forward: statistics + extrema
// =================================================================================================
template <int CI>
__global__ void __launch_bounds__(kPWWarps * 32) pwmlp_fwd_kernel(const PwArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * 32 * CI;
  const int chunkC = min(32 * CI, a.Cop - c0);
  const uint32_t row_bytes = (uint32_t)chunkC * 4u;
  const PwSmem L = pw_smem(a.K, 32 * CI);
  float* s_stage = reinterpret_cast<float*>(smem + L.stage_off + (size_t)warp * kPWStageBytes);
  float* s_a = reinterpret_cast<float*>(smem + L.a_off) + (size_t)warp * 32 * CI;
  float4* s_dp = reinterpret_cast<float4*>(smem + L.dp_off) + (size_t)warp * a.K;
  int* s_idx = reinterpret_cast<int*>(smem + L.idx_off) + (size_t)warp * a.K;
  float* s_max = reinterpret_cast<float*>(smem + L.t0_off);
  float* s_min = reinterpret_cast<float*>(smem + L.t1_off);
  unsigned short* s_arg = reinterpret_cast<unsigned short*>(smem + L.arg_off);
  float* s_red = reinterpret_cast<float*>(smem + L.red_off);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off) + warp;

  const int tiles_per_cloud = (a.M + kPWTile - 1) / kPWTile;
  const int b = blockIdx.x / tiles_per_cloud;
  const int q0 = (blockIdx.x % tiles_per_cloud) * kPWTile;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  uint32_t phase = 0;

  float wpx[CI], wpy[CI], wpz[CI], s1[CI], s2[CI];
#pragma unroll
  for (int i = 0; i < CI; ++i) {
    const int c = c0 + lane + 32 * i;
    const bool ok = c < a.Cout;
    wpx[i] = ok ? a.wp[c * 3 + 0] : 0.f;
    wpy[i] = ok ? a.wp[c * 3 + 1] : 0.f;
    wpz[i] = ok ? a.wp[c * 3 + 2] : 0.f;
    s1[i] = s2[i] = 0.f;
  }
  const float* ab = a.ab_pm + (size_t)b * a.N * 2 * a.Cop;
  const float* sxyz = a.support_xyz + (size_t)b * a.N * 3;

  for (int ql = warp; ql < kPWTile; ql += kPWWarps) {
    const int q = q0 + ql;
    float vmax[CI], vmin[CI];
    int kmax[CI], kmin[CI];
#pragma unroll
    for (int i = 0; i < CI; ++i) {
      vmax[i] = -INFINITY;
      vmin[i] = INFINITY;
      kmax[i] = kmin[i] = 0;
    }
    if (q < a.M) {
      const size_t gq = (size_t)b * a.M + q;
      const float qx = a.query_xyz[gq * 3 + 0], qy = a.query_xyz[gq * 3 + 1], qz = a.query_xyz[gq * 3 + 2];
      for (int k = lane; k < a.K; k += 32) {
        const int j = a.idx[gq * a.K + k];
        s_idx[k] = j;
        const float dx = __fmul_rn(__fsub_rn(sxyz[j * 3 + 0], qx), a.inv_radius);
        const float dy = __fmul_rn(__fsub_rn(sxyz[j * 3 + 1], qy), a.inv_radius);
        const float dz = __fmul_rn(__fsub_rn(sxyz[j * 3 + 2], qz), a.inv_radius);
        s_dp[k] = make_float4(dx, dy, dz, 0.f);
      }
      __syncwarp();
      for (int k0 = 0; k0 < a.K; k0 += a.rows_per_stage) {
        const int rows = min(a.rows_per_stage, a.K - k0);
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(rows + (k0 == 0 ? 1 : 0)) * row_bytes);
        __syncwarp();
        if (k0 == 0 && lane == 0)  // centre row: A[j_0], slot 0 = nearest neighbour (:290)
          bulk_g2s(s_a, ab + (size_t)s_idx[0] * 2 * a.Cop + c0, row_bytes, bar);
        for (int kk = lane; kk < rows; kk += 32)
          bulk_g2s(s_stage + (size_t)kk * chunkC, ab + (size_t)s_idx[k0 + kk] * 2 * a.Cop + a.Cop + c0, row_bytes, bar);
        mbar_wait(bar, phase);
        phase ^= 1u;
        float av[CI];
#pragma unroll
        for (int i = 0; i < CI; ++i) av[i] = (lane + 32 * i < chunkC) ? s_a[lane + 32 * i] : 0.f;
        for (int kk = 0; kk < rows; ++kk) {
          const float4 dp = s_dp[k0 + kk];
          const float* row = s_stage + (size_t)kk * chunkC;
#pragma unroll
          for (int i = 0; i < CI; ++i) {
            if (lane + 32 * i < chunkC) {
              const float y = fmaf(wpz[i], dp.z, fmaf(wpy[i], dp.y, fmaf(wpx[i], dp.x, av[i] + row[lane + 32 * i])));
              s1[i] += y;
              s2[i] = fmaf(y, y, s2[i]);
              if (y > vmax[i]) { vmax[i] = y; kmax[i] = k0 + kk; }  // first occurrence wins, as max_pool2d
              if (y < vmin[i]) { vmin[i] = y; kmin[i] = k0 + kk; }
            }
          }
        }
        __syncwarp();
      }
    }
#pragma unroll
    for (int i = 0; i < CI; ++i) {
      s_max[(size_t)(lane + 32 * i) * (kPWTile + 1) + ql] = vmax[i];
      s_min[(size_t)(lane + 32 * i) * (kPWTile + 1) + ql] = vmin[i];
      s_arg[(size_t)ql * 32 * CI + lane + 32 * i] = (unsigned short)(kmax[i] | (kmin[i] << 8));
    }
  }
#pragma unroll
  for (int i = 0; i < CI; ++i) {
    s_red[((size_t)warp * 3 + 0) * 32 * CI + lane + 32 * i] = s1[i];
    s_red[((size_t)warp * 3 + 1) * 32 * CI + lane + 32 * i] = s2[i];
  }
  __syncthreads();
  const int q = q0 + lane;
  for (int cl = warp; cl < 32 * CI; cl += kPWWarps) {
    const int c = c0 + cl;
    if (c >= a.Cout) break;
    if (q < a.M) {
      a.ymax[((size_t)b * a.Cout + c) * a.M + q] = s_max[(size_t)cl * (kPWTile + 1) + lane];
      a.ymin[((size_t)b * a.Cout + c) * a.M + q] = s_min[(size_t)cl * (kPWTile + 1) + lane];
    }
  }
  const int nq = min(kPWTile, a.M - q0);
  for (int e = threadIdx.x; e < nq * chunkC; e += blockDim.x) {
    const int ql = e / chunkC, cl = e % chunkC;
    a.arg[((size_t)b * a.M + q0 + ql) * a.Cop + c0 + cl] = s_arg[(size_t)ql * 32 * CI + cl];
  }
  for (int cl = threadIdx.x; cl < 32 * CI; cl += blockDim.x) {
    const int c = c0 + cl;
    if (c >= a.Cout) continue;
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int w = 0; w < kPWWarps; ++w) {
      t1 += s_red[((size_t)w * 3 + 0) * 32 * CI + cl];
      t2 += s_red[((size_t)w * 3 + 1) * 32 * CI + cl];
    }
    a.partial[((size_t)blockIdx.x * 2 + 0) * a.Cout + c] = t1;
    a.partial[((size_t)blockIdx.x * 2 + 1) * a.Cout + c] = t2;
  }
}

// out[b,o,q] = relu(sc*(sc>=0 ? ymax : ymin) + sh)   -- max over K through the monotone BN+ReLU
__global__ void __launch_bounds__(256) pwmlp_out_kernel(const float* __restrict__ ymax, const float* __restrict__ ymin,
                                                        const float* __restrict__ stats,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int C, int M,
                                                        float* __restrict__ out) {
  const int row = blockIdx.x;
  const int c = row % C;
  const float mean = stats[c], sc = stats[C + c] * gamma[c], sh = beta[c];
  const float* src = (sc >= 0.f ? ymax : ymin) + (size_t)row * M;
  float* dst = out + (size_t)row * M;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < M; i += gridDim.y * blockDim.x) {
    const float v = __fmaf_rn(__fsub_rn(src[i], mean), sc, sh);
    dst[i] = v > 0.f ? v : 0.f;
  }
}

// =================================================================================================
// backward
// =================================================================================================
// per tile & channel: sum dz, sum dz*yhat at the selected slot (BatchNorm2d backward needs them over B*M*K;
// dz is non-zero only at the arg-max / arg-min slot)
__global__ void __launch_bounds__(256) pwmlp_bwd_stats_kernel(const float* __restrict__ grad_out,
                                                              const float* __restrict__ out,
                                                              const float* __restrict__ ymax,
                                                              const float* __restrict__ ymin,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ gamma, int C, int M,
                                                              float* __restrict__ partial) {
  const int tiles_per_cloud = (M + kPWTile - 1) / kPWTile;
  const int b = blockIdx.x / tiles_per_cloud;
  const int q = (blockIdx.x % tiles_per_cloud) * kPWTile + (threadIdx.x & 31);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < C; c += 8) {
    float dz = 0.f, dzy = 0.f;
    if (q < M) {
      const size_t o = ((size_t)b * C + c) * M + q;
      const float sc = stats[C + c] * gamma[c];
      const float ysel = sc >= 0.f ? ymax[o] : ymin[o];
      dz = out[o] > 0.f ? grad_out[o] : 0.f;
      dzy = dz * ((ysel - stats[c]) * stats[C + c]);
    }
    const float t1 = warp_sum(dz), t2 = warp_sum(dzy);
    if (lane == 0) {
      partial[((size_t)blockIdx.x * 2 + 0) * C + c] = t2;  // -> dgamma = sum dz*yhat
      partial[((size_t)blockIdx.x * 2 + 1) * C + c] = t1;  // -> dbeta  = sum dz
    }
  }
}

template <int CI>
__global__ void __launch_bounds__(kPWWarps * 32) pwmlp_bwd_kernel(const PwArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * 32 * CI;
  const int chunkC = min(32 * CI, a.Cop - c0);
  const uint32_t row_bytes = (uint32_t)chunkC * 4u;
  const PwSmem L = pw_smem(a.K, 32 * CI);
  float* s_stage = reinterpret_cast<float*>(smem + L.stage_off + (size_t)warp * kPWStageBytes);
  float* s_a = reinterpret_cast<float*>(smem + L.a_off) + (size_t)warp * 32 * CI;
  float4* s_dp = reinterpret_cast<float4*>(smem + L.dp_off) + (size_t)warp * a.K;
  int* s_idx = reinterpret_cast<int*>(smem + L.idx_off) + (size_t)warp * a.K;
  float* s_dz = reinterpret_cast<float*>(smem + L.t0_off);  // [chunk][tile+1]: sc * dz
  float* s_red = reinterpret_cast<float*>(smem + L.red_off);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off) + warp;

  const int tiles_per_cloud = (a.M + kPWTile - 1) / kPWTile;
  const int b = blockIdx.x / tiles_per_cloud;
  const int q0 = (blockIdx.x % tiles_per_cloud) * kPWTile;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  uint32_t phase = 0;

  // tile of sc*dz, coalesced from the channel-major gradient
  {
    const int q = q0 + lane;
    for (int cl = warp; cl < 32 * CI; cl += kPWWarps) {
      const int c = c0 + cl;
      float v = 0.f;
      if (c < a.Cout && q < a.M) {
        const size_t o = ((size_t)b * a.Cout + c) * a.M + q;
        v = a.out[o] > 0.f ? a.grad_out[o] * (a.stats[a.Cout + c] * a.gamma[c]) : 0.f;
      }
      s_dz[(size_t)cl * (kPWTile + 1) + lane] = v;
    }
  }
  __syncthreads();

  float wpx[CI], wpy[CI], wpz[CI], c1[CI], c2[CI], mean[CI], dwx[CI], dwy[CI], dwz[CI];
  bool use_max[CI];
#pragma unroll
  for (int i = 0; i < CI; ++i) {
    const int c = c0 + lane + 32 * i;
    const bool ok = c < a.Cout;
    wpx[i] = ok ? a.wp[c * 3 + 0] : 0.f;
    wpy[i] = ok ? a.wp[c * 3 + 1] : 0.f;
    wpz[i] = ok ? a.wp[c * 3 + 2] : 0.f;
    const float invstd = ok ? a.stats[a.Cout + c] : 0.f;
    const float sc = ok ? invstd * a.gamma[c] : 0.f;
    mean[i] = ok ? a.stats[c] : 0.f;
    // dy = sc*[k==k*]*dz - sc*s1/P - sc*invstd*(y-mean)*s2/P     (BatchNorm2d backward, P = B*M*K)
    c1[i] = ok ? sc * a.dgb[a.Cout + c] * a.inv_count : 0.f;
    c2[i] = ok ? sc * invstd * a.dgb[c] * a.inv_count : 0.f;
    use_max[i] = sc >= 0.f;
    dwx[i] = dwy[i] = dwz[i] = 0.f;
  }
  const float* ab = a.ab_pm + (size_t)b * a.N * 2 * a.Cop;
  float* gab = a.grad_ab_pm + (size_t)b * a.N * 2 * a.Cop;
  const float* sxyz = a.support_xyz + (size_t)b * a.N * 3;

  for (int ql = warp; ql < kPWTile; ql += kPWWarps) {
    const int q = q0 + ql;
    if (q >= a.M) continue;
    const size_t gq = (size_t)b * a.M + q;
    const float qx = a.query_xyz[gq * 3 + 0], qy = a.query_xyz[gq * 3 + 1], qz = a.query_xyz[gq * 3 + 2];
    for (int k = lane; k < a.K; k += 32) {
      const int j = a.idx[gq * a.K + k];
      s_idx[k] = j;
      const float dx = __fmul_rn(__fsub_rn(sxyz[j * 3 + 0], qx), a.inv_radius);
      const float dy = __fmul_rn(__fsub_rn(sxyz[j * 3 + 1], qy), a.inv_radius);
      const float dz = __fmul_rn(__fsub_rn(sxyz[j * 3 + 2], qz), a.inv_radius);
      s_dp[k] = make_float4(dx, dy, dz, 0.f);
    }
    __syncwarp();
    float dzs[CI], da[CI];
    int kstar[CI];
#pragma unroll
    for (int i = 0; i < CI; ++i) {
      dzs[i] = s_dz[(size_t)(lane + 32 * i) * (kPWTile + 1) + ql];
      da[i] = 0.f;
      kstar[i] = -1;
      if (lane + 32 * i < chunkC) {
        const unsigned short ar = a.arg[gq * a.Cop + c0 + lane + 32 * i];
        kstar[i] = use_max[i] ? (ar & 0xff) : (ar >> 8);
      }
    }
    for (int k0 = 0; k0 < a.K; k0 += a.rows_per_stage) {
      const int rows = min(a.rows_per_stage, a.K - k0);
      if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(rows + (k0 == 0 ? 1 : 0)) * row_bytes);
      __syncwarp();
      if (k0 == 0 && lane == 0) bulk_g2s(s_a, ab + (size_t)s_idx[0] * 2 * a.Cop + c0, row_bytes, bar);
      for (int kk = lane; kk < rows; kk += 32)
        bulk_g2s(s_stage + (size_t)kk * chunkC, ab + (size_t)s_idx[k0 + kk] * 2 * a.Cop + a.Cop + c0, row_bytes, bar);
      mbar_wait(bar, phase);
      phase ^= 1u;
      float av[CI];
#pragma unroll
      for (int i = 0; i < CI; ++i) av[i] = (lane + 32 * i < chunkC) ? s_a[lane + 32 * i] : 0.f;
      for (int kk = 0; kk < rows; ++kk) {
        const float4 dp = s_dp[k0 + kk];
        const float* row = s_stage + (size_t)kk * chunkC;
        float* grow = gab + (size_t)s_idx[k0 + kk] * 2 * a.Cop + a.Cop + c0;
#pragma unroll
        for (int i = 0; i < CI; ++i) {
          if (c0 + lane + 32 * i < a.Cout) {
            const float y = fmaf(wpz[i], dp.z, fmaf(wpy[i], dp.y, fmaf(wpx[i], dp.x, av[i] + row[lane + 32 * i])));
            const float dy = ((k0 + kk) == kstar[i] ? dzs[i] : 0.f) - c1[i] - c2[i] * (y - mean[i]);
            da[i] += dy;
            dwx[i] = fmaf(dy, dp.x, dwx[i]);
            dwy[i] = fmaf(dy, dp.y, dwy[i]);
            dwz[i] = fmaf(dy, dp.z, dwz[i]);
            atomicAdd(grow + lane + 32 * i, dy);  // d/dBv[j_k]  (red.global.add.f32, coalesced per row)
          }
        }
      }
      __syncwarp();
    }
    float* garow = gab + (size_t)s_idx[0] * 2 * a.Cop + c0;  // d/dA[j_0] += sum_k dy
#pragma unroll
    for (int i = 0; i < CI; ++i)
      if (c0 + lane + 32 * i < a.Cout) atomicAdd(garow + lane + 32 * i, da[i]);
    __syncwarp();
  }
  // d/dWp partials: fixed-order reduction over the CTA's warps
#pragma unroll
  for (int i = 0; i < CI; ++i) {
    s_red[((size_t)warp * 3 + 0) * 32 * CI + lane + 32 * i] = dwx[i];
    s_red[((size_t)warp * 3 + 1) * 32 * CI + lane + 32 * i] = dwy[i];
    s_red[((size_t)warp * 3 + 2) * 32 * CI + lane + 32 * i] = dwz[i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 3 * 32 * CI; e += blockDim.x) {
    const int s = e / (32 * CI), cl = e % (32 * CI);
    const int c = c0 + cl;
    if (c >= a.Cout) continue;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kPWWarps; ++w) t += s_red[((size_t)w * 3 + s) * 32 * CI + cl];
    a.partial[((size_t)blockIdx.x * 3 + s) * a.Cout + c] = t;
  }
}

template <int CI>
static int launch_pw_fwd(const PwArgs& a, cudaStream_t stream) {
  const PwSmem L = pw_smem(a.K, 32 * CI);
  cudaFuncSetAttribute(pwmlp_fwd_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  dim3 grid(a.ntiles, ceil_div(a.Cop, 32 * CI));
  pwmlp_fwd_kernel<CI><<<grid, kPWWarps * 32, L.total, stream>>>(a); CL3D_LAUNCHED(1);
  return check_launch("pwmlp_fwd_kernel");
}
template <int CI>
static int launch_pw_bwd(const PwArgs& a, cudaStream_t stream) {
  const PwSmem L = pw_smem(a.K, 32 * CI);
  cudaFuncSetAttribute(pwmlp_bwd_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  dim3 grid(a.ntiles, ceil_div(a.Cop, 32 * CI));
  pwmlp_bwd_kernel<CI><<<grid, kPWWarps * 32, L.total, stream>>>(a); CL3D_LAUNCHED(1);
  return check_launch("pwmlp_bwd_kernel");
}

static int pw_ci(int Cop) {
  int ci = ceil_div(Cop, 32);
  return ci > kPWMaxCI ? kPWMaxCI : ci;
}
static void pw_rows(PwArgs& a, int ci) {
  const int chunk = 32 * ci < a.Cop ? 32 * ci : a.Cop;
  a.rows_per_stage = kPWStageBytes / (chunk * 4);
  if (a.rows_per_stage > a.K) a.rows_per_stage = a.K;
}

}  // namespace cl3d

using namespace cl3d;

extern "C" int cl3d_pwmlp_fwd_stats(const float* ab_pm, const float* wp, const float* query_xyz,
                                    const float* support_xyz, const int* idx, int B, int N, int M, int K, int Cout,
                                    float radius, float* ymax, float* ymin, unsigned short* arg, float* bn_partial,
                                    cl3d_stream_t stream_) {
  CL3D_REQUIRE(ab_pm && wp && query_xyz && support_xyz && idx && ymax && ymin && arg && bn_partial,
               "cl3d_pwmlp_fwd_stats: null pointer");
  CL3D_REQUIRE(B >= 0 && N >= 1 && M >= 1 && K >= 1 && K <= 255 && Cout >= 1, "cl3d_pwmlp_fwd_stats: bad sizes (K <= 255)");
  if (B == 0) return CL3D_OK;
  PwArgs a = {};
  a.ab_pm = ab_pm; a.wp = wp; a.query_xyz = query_xyz; a.support_xyz = support_xyz; a.idx = idx;
  a.ymax = ymax; a.ymin = ymin; a.arg = arg; a.partial = bn_partial;
  a.B = B; a.N = N; a.M = M; a.K = K; a.Cout = Cout; a.Cop = padded_channels(Cout);
  a.inv_radius = 1.0f / radius;
  a.ntiles = B * ceil_div(M, kPWTile);
  const int ci = pw_ci(a.Cop);
  pw_rows(a, ci);
  switch (ci) {
    case 1: return launch_pw_fwd<1>(a, (cudaStream_t)stream_);
    case 2: return launch_pw_fwd<2>(a, (cudaStream_t)stream_);
    case 3: return launch_pw_fwd<3>(a, (cudaStream_t)stream_);
    default: return launch_pw_fwd<4>(a, (cudaStream_t)stream_);
  }
}

extern "C" int cl3d_pwmlp_fwd_out(const float* ymax, const float* ymin, const float* save_stats, const float* gamma,
                                  const float* beta, int B, int M, int Cout, float* out, cl3d_stream_t stream_) {
  CL3D_REQUIRE(ymax && ymin && save_stats && gamma && beta && out && B >= 0 && M >= 1 && Cout >= 1,
               "cl3d_pwmlp_fwd_out: bad arguments");
  if (B == 0) return CL3D_OK;
  int chunks = ceil_div(M, 1024);
  chunks = chunks < 1 ? 1 : (chunks > 65535 ? 65535 : chunks);
  pwmlp_out_kernel<<<dim3(B * Cout, chunks), 256, 0, (cudaStream_t)stream_>>>(ymax, ymin, save_stats, gamma, beta,
                                                                             Cout, M, out); CL3D_LAUNCHED(1);
  return check_launch("pwmlp_out_kernel");
}

extern "C" int cl3d_pwmlp_bwd(const float* grad_out, const float* out, const float* ab_pm, const float* wp,
                              const float* query_xyz, const float* support_xyz, const int* idx, const float* ymax,
                              const float* ymin, const unsigned short* arg, const float* save_stats,
                              const float* gamma, int B, int N, int M, int K, int Cout, float radius,
                              float* partial /* (ntiles, 3, Cout) scratch */, float* dgamma_dbeta /* (2,Cout) */,
                              float* grad_ab_pm /* (B,N,2*Cop), zero-filled here */, float* grad_wp /* (3,Cout) */,
                              cl3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CL3D_REQUIRE(grad_out && out && ab_pm && wp && query_xyz && support_xyz && idx && ymax && ymin && arg &&
                   save_stats && gamma && partial && dgamma_dbeta && grad_ab_pm && grad_wp,
               "cl3d_pwmlp_bwd: null pointer");
  CL3D_REQUIRE(B >= 0 && N >= 1 && M >= 1 && K >= 1 && K <= 255 && Cout >= 1, "cl3d_pwmlp_bwd: bad sizes");
  if (B == 0) return CL3D_OK;
  const int ntiles = B * ceil_div(M, kPWTile);
  const int Cop = padded_channels(Cout);
  pwmlp_bwd_stats_kernel<<<ntiles, 256, 0, stream>>>(grad_out, out, ymax, ymin, save_stats, gamma, Cout, M, partial); CL3D_LAUNCHED(1);
  // dgamma_dbeta[0] = sum dz*yhat (dgamma), [1] = sum dz (dbeta)
  int rc = cl3d_reduce_partials(partial, ntiles, 2 * Cout, dgamma_dbeta, stream_);
  if (rc) return rc;
  cudaMemsetAsync(grad_ab_pm, 0, sizeof(float) * (size_t)B * N * 2 * Cop, stream);
  PwArgs a = {};
  a.ab_pm = ab_pm; a.wp = wp; a.query_xyz = query_xyz; a.support_xyz = support_xyz; a.idx = idx;
  a.arg = const_cast<unsigned short*>(arg);
  a.partial = partial;
  a.grad_out = grad_out; a.out = out; a.stats = save_stats; a.gamma = gamma; a.dgb = dgamma_dbeta;
  a.grad_ab_pm = grad_ab_pm;
  a.B = B; a.N = N; a.M = M; a.K = K; a.Cout = Cout; a.Cop = Cop;
  a.inv_radius = 1.0f / radius;
  a.inv_count = 1.0f / (float)((double)B * M * K);
  a.ntiles = ntiles;
  const int ci = pw_ci(Cop);
  pw_rows(a, ci);
  switch (ci) {
    case 1: rc = launch_pw_bwd<1>(a, stream); break;
    case 2: rc = launch_pw_bwd<2>(a, stream); break;
    case 3: rc = launch_pw_bwd<3>(a, stream); break;
    default: rc = launch_pw_bwd<4>(a, stream); break;
  }
  if (rc) return rc;
  return cl3d_reduce_partials(partial, ntiles, 3 * Cout, grad_wp, stream_);
}
