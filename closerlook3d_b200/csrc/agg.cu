// agg.cu -- fused local aggregation, forward and backward, for the per-(channel, neighbour)-weight families
//           PosPool(xyz | sin_cos), AdaptiveWeight(dp) and PseudoGrid (sm_100a).
//
// Replaces, per LocalAggregation call of the reference
//   /root/reference/pytorch/ops/pt_custom_ops/pt_utils.py:121-144            (2x group_points, subtract, normalise)
//   /root/reference/pytorch/models/local_aggregation_operators.py:65-105      (PosPool transform + reduction)
//   .../local_aggregation_operators.py:188-217                                (AdaptiveWeight)
//   .../local_aggregation_operators.py:384-419                                (PseudoGrid)
// none of which exists as a fused kernel there: the reference materialises (B,C,M,K) tensors ~10 times.
//
// Forward  (one warp per query, 32 consecutive queries per CTA):
//   lanes load the query's neighbour indices, gather the neighbours' xyz, build dp = (s - q) * (1/r);
//   every lane then issues one bulk async copy (cp.async.bulk, TMA non-tensor form) per neighbour ROW of
//   the point-major feature matrix into shared memory, completion on an mbarrier; the warp multiplies by
//   the family weight w_c(dp_k) and reduces over K in registers (lane = channel), so only the aggregated
//   (B,C,M) tensor is written -- transposed through shared memory into the reference's channel-major
//   layout, together with per-tile BatchNorm partial sums.
// Backward (gather form, one warp per SUPPORT point over its CSR list of (query, slot) references):
//   rows of d(loss)/d(agg) are bulk-copied the same way; no float atomics on activations; parameter
//   gradients are accumulated per CTA and reduced afterwards in a fixed order.
#include "common.cuh"

namespace cl3d {

constexpr int kAggWarps = 8;
constexpr int kTile = 32;          // queries (fwd) / support points (bwd) per CTA tile
constexpr int kStageBytes = 8192;  // bulk-copy staging per warp
constexpr int kMaxKP = 16;         // PseudoGrid kernel points (reference default 15)
constexpr int kMaxCI = 6;          // channel chunk = 32*CI <= 192 channels per CTA

struct AggArgs {
  const float* feat_pm;      // (B,N,Cp)   fwd: features; bwd: features (for parameter gradients)
  const float* g_pm;         // (B,M,Cp)   bwd only
  const float* query_xyz;    // (B,M,3)
  const float* support_xyz;  // (B,N,3)
  const int* idx;            // (B,M,K)    fwd only
  const int* ncount;         // (B,M)
  const int* csr_off;        // (B,N+1)    bwd only
  const int* csr_ent;        // (B,M*K)    bwd only
  const float* p0;           // family parameter 0 (see cl3d.h)
  const float* p1;           // family parameter 1
  float* out;                // fwd: agg (B,C,M); bwd: grad_feat (B,C,N)
  float* partial;            // fwd: bn partial (ntiles,2,C); bwd: param-grad partial (gridDim.x, P)
  int B, N, M, K, C, Cp;
  int reduction, normalize, shared, nkp, influence;
  float inv_radius, extent, inv_extent;
  int rows_per_stage;        // bulk-copy rows per stage for this channel chunk
  int ntiles;
};

// ---------------------------------------------------------------------------------------------
// family weights
// ---------------------------------------------------------------------------------------------
template <int FAM, int CI>
struct LaneParams {  // per-lane, per owned channel constants
  int axis[CI];      // XYZ / SINCOS: which coordinate
  float a[CI];       // SINCOS: dim_mat value ; ADAPTIVE: Wx
  float b[CI];       // ADAPTIVE: Wy
  float c[CI];       // ADAPTIVE: Wz
  float d[CI];       // ADAPTIVE: bias
  int is_cos[CI];
};

template <int FAM, int CI>
__device__ __forceinline__ void load_lane_params(LaneParams<FAM, CI>& lp, const AggArgs& a, int c0, int lane) {
#pragma unroll
  for (int i = 0; i < CI; ++i) {
    const int c = c0 + lane + 32 * i;
    lp.axis[i] = 0;
    lp.a[i] = 1.f;
    lp.b[i] = lp.c[i] = lp.d[i] = 0.f;
    lp.is_cos[i] = 0;
    if (c >= a.C) continue;
    if (FAM == CL3D_FAM_POSPOOL_XYZ) {
      lp.axis[i] = c % 3;  // view(B, C//3, 3, ...) : local_aggregation_operators.py:67
    } else if (FAM == CL3D_FAM_POSPOOL_SINCOS) {
      const int F = a.C / 6;  // channel = axis*2F + t ; t<F sin, t>=F cos  (:70-83)
      const int t = c % (2 * F);
      lp.axis[i] = c / (2 * F);
      lp.is_cos[i] = t >= F;
      lp.a[i] = a.p0[t % F];
    } else if (FAM == CL3D_FAM_ADAPTIVE_DP) {
      const int g = c / a.shared;  // :194-197 channel c uses weight row c // S
      lp.a[i] = a.p0[g * 3 + 0];
      lp.b[i] = a.p0[g * 3 + 1];
      lp.c[i] = a.p0[g * 3 + 2];
      lp.d[i] = a.p1[g];
    }
  }
}

// weight of channel slot i for relative position dp (float4: x,y,z,scale)
template <int FAM, int CI>
__device__ __forceinline__ float family_weight(const LaneParams<FAM, CI>& lp, int i, const float4& dp) {
  if (FAM == CL3D_FAM_POSPOOL_XYZ) {
    return lp.axis[i] == 0 ? dp.x : (lp.axis[i] == 1 ? dp.y : dp.z);
  } else if (FAM == CL3D_FAM_POSPOOL_SINCOS) {
    const float p = lp.axis[i] == 0 ? dp.x : (lp.axis[i] == 1 ? dp.y : dp.z);
    const float arg = __fdiv_rn(__fmul_rn(100.f, p), lp.a[i]);  // torch.div(alpha * dp, dim_mat) :75-77
    return lp.is_cos[i] ? cosf(arg) : sinf(arg);
  } else if (FAM == CL3D_FAM_ADAPTIVE_DP) {
    return fmaf(lp.c[i], dp.z, fmaf(lp.b[i], dp.y, fmaf(lp.a[i], dp.x, lp.d[i])));  // 1x1 conv 3 -> C/S, bias
  }
  return 0.f;
}

// PseudoGrid influence of kernel point kp on relative position dp (:385-403), mask applied by caller
__device__ __forceinline__ float pg_influence(float dx, float dy, float dz, const float* kp, int influence,
                                              float extent) {
  if (influence == 1) return 1.f;  // 'constant'
  const float ex = dx - kp[0], ey = dy - kp[1], ez = dz - kp[2];
  const float sq = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
  const float h = 1.f - __fdiv_rn(sqrtf(sq), extent);  // clamp(1 - sqrt(sq)/extent, min=0)  :397
  return h > 0.f ? h : 0.f;
}

// shared-memory carve-up (per CTA)
struct SmemLayout {
  size_t stage_off, dp_off, idx_off, h_off, out_off, bar_off, red_off, total;
};
__host__ __device__ inline SmemLayout smem_layout(int K_or_rows, int chunkC, bool pseudogrid, int red_floats) {
  SmemLayout L;
  size_t o = 0;
  L.stage_off = o;
  o += (size_t)kAggWarps * kStageBytes;
  L.dp_off = o;
  o += (size_t)kAggWarps * K_or_rows * sizeof(float4);
  L.idx_off = o;
  o += (size_t)kAggWarps * K_or_rows * sizeof(int);
  o = align_up(o, 16);
  L.h_off = o;
  o += pseudogrid ? (size_t)kAggWarps * K_or_rows * kMaxKP * sizeof(float) : 0;
  L.out_off = o;
  o += (size_t)chunkC * (kTile + 1) * sizeof(float);
  o = align_up(o, 16);
  L.red_off = o;
  o += (size_t)red_floats * sizeof(float);
  o = align_up(o, 16);
  L.bar_off = o;
  o += (size_t)kAggWarps * sizeof(uint64_t);
  L.total = o;
  return L;
}

// =================================================================================================
// forward
// =================================================================================================
template <int FAM, int CI>
__global__ void __launch_bounds__(kAggWarps * 32) agg_fwd_kernel(const AggArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * 32 * CI;
  const int chunkC = min(32 * CI, a.Cp - c0);  // multiple of 8
  const uint32_t row_bytes = (uint32_t)chunkC * 4u;
  const SmemLayout L = smem_layout(a.K, 32 * CI, FAM == CL3D_FAM_PSEUDOGRID, 0);
  float* s_stage = reinterpret_cast<float*>(smem + L.stage_off + (size_t)warp * kStageBytes);
  float4* s_dp = reinterpret_cast<float4*>(smem + L.dp_off) + (size_t)warp * a.K;
  int* s_idx = reinterpret_cast<int*>(smem + L.idx_off) + (size_t)warp * a.K;
  float* s_h = reinterpret_cast<float*>(smem + L.h_off) + (size_t)warp * a.K * kMaxKP;
  float* s_out = reinterpret_cast<float*>(smem + L.out_off);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off) + warp;

  const int tiles_per_cloud = (a.M + kTile - 1) / kTile;
  const int b = blockIdx.x / tiles_per_cloud;
  const int q0 = (blockIdx.x % tiles_per_cloud) * kTile;

  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  uint32_t phase = 0;

  LaneParams<FAM, CI> lp;
  load_lane_params<FAM, CI>(lp, a, c0, lane);
  // PseudoGrid: kernel weights of my channels stay in registers, kernel points in smem via s_h prologue
  float wk[FAM == CL3D_FAM_PSEUDOGRID ? kMaxKP : 1][CI];
  if (FAM == CL3D_FAM_PSEUDOGRID) {
#pragma unroll
    for (int kp = 0; kp < kMaxKP; ++kp)
#pragma unroll
      for (int i = 0; i < CI; ++i) {
        const int c = c0 + lane + 32 * i;
        wk[kp][i] = (kp < a.nkp && c < a.C) ? a.p1[(size_t)kp * a.C + c] : 0.f;
      }
  }

  const float* feat = a.feat_pm + (size_t)b * a.N * a.Cp + c0;
  const float* sxyz = a.support_xyz + (size_t)b * a.N * 3;

  for (int ql = warp; ql < kTile; ql += kAggWarps) {
    const int q = q0 + ql;
    float res[CI];
#pragma unroll
    for (int i = 0; i < CI; ++i) res[i] = 0.f;
    if (q < a.M) {
      const size_t gq = (size_t)b * a.M + q;
      const int nrows = a.ncount[gq];
      const float qx = a.query_xyz[gq * 3 + 0], qy = a.query_xyz[gq * 3 + 1], qz = a.query_xyz[gq * 3 + 2];
      // ---- prologue: indices, relative positions (pt_utils.py:127-129), PseudoGrid influences
      for (int k = lane; k < nrows; k += 32) {
        const int j = a.idx[gq * a.K + k];
        s_idx[k] = j;
        float dx = __fsub_rn(sxyz[j * 3 + 0], qx), dy = __fsub_rn(sxyz[j * 3 + 1], qy),
              dz = __fsub_rn(sxyz[j * 3 + 2], qz);
        if (a.normalize) {  // torch's CUDA `tensor /= python_scalar` multiplies by the fp32 reciprocal
          dx = __fmul_rn(dx, a.inv_radius);
          dy = __fmul_rn(dy, a.inv_radius);
          dz = __fmul_rn(dz, a.inv_radius);
        }
        s_dp[k] = make_float4(dx, dy, dz, 0.f);
        if (FAM == CL3D_FAM_PSEUDOGRID) {
          for (int kp = 0; kp < a.nkp; ++kp)
            s_h[k * kMaxKP + kp] = pg_influence(dx, dy, dz, a.p0 + kp * 3, a.influence, a.extent);
          for (int kp = a.nkp; kp < kMaxKP; ++kp) s_h[k * kMaxKP + kp] = 0.f;
        }
      }
      __syncwarp();
      float acc[FAM == CL3D_FAM_PSEUDOGRID ? kMaxKP : 1][CI];
#pragma unroll
      for (int kp = 0; kp < (FAM == CL3D_FAM_PSEUDOGRID ? kMaxKP : 1); ++kp)
#pragma unroll
        for (int i = 0; i < CI; ++i) acc[kp][i] = 0.f;

      for (int k0 = 0; k0 < nrows; k0 += a.rows_per_stage) {
        const int rows = min(a.rows_per_stage, nrows - k0);
        // ---- stage `rows` neighbour rows through TMA bulk copies (one per lane)
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)rows * row_bytes);
        __syncwarp();
        for (int kk = lane; kk < rows; kk += 32)
          bulk_g2s(s_stage + (size_t)kk * chunkC, feat + (size_t)s_idx[k0 + kk] * a.Cp, row_bytes, bar);
        mbar_wait(bar, phase);
        phase ^= 1u;
        // ---- transform + reduce over the staged rows
        for (int kk = 0; kk < rows; ++kk) {
          const float4 dp = s_dp[k0 + kk];
          const float* row = s_stage + (size_t)kk * chunkC;
          if (FAM == CL3D_FAM_PSEUDOGRID) {
            float v[CI];
#pragma unroll
            for (int i = 0; i < CI; ++i) v[i] = (lane + 32 * i < chunkC) ? row[lane + 32 * i] : 0.f;
            const float* hk = s_h + (size_t)(k0 + kk) * kMaxKP;
#pragma unroll
            for (int kp = 0; kp < kMaxKP; ++kp) {
              const float h = hk[kp];  // entries >= nkp are never read as non-zero: wk is 0 there
#pragma unroll
              for (int i = 0; i < CI; ++i) acc[kp][i] = fmaf(h, v[i], acc[kp][i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < CI; ++i) {
              if (lane + 32 * i < chunkC) {
                const float w = family_weight<FAM, CI>(lp, i, dp);
                acc[0][i] = fmaf(row[lane + 32 * i], w, acc[0][i]);
              }
            }
          }
        }
        __syncwarp();  // all lanes done reading the stage before it is refilled
      }
      if (FAM == CL3D_FAM_PSEUDOGRID) {
#pragma unroll
        for (int i = 0; i < CI; ++i) {
          float r = 0.f;
#pragma unroll
          for (int kp = 0; kp < kMaxKP; ++kp) r = fmaf(acc[kp][i], wk[kp][i], r);  // sum_k' Wk[k',c] * t[k',c]  :418-419
          res[i] = r;
        }
      } else {
#pragma unroll
        for (int i = 0; i < CI; ++i)
          res[i] = a.reduction == CL3D_REDUCE_AVG ? __fdiv_rn(acc[0][i], (float)nrows) : acc[0][i];  // :96-98
      }
    }
#pragma unroll
    for (int i = 0; i < CI; ++i) s_out[(size_t)(lane + 32 * i) * (kTile + 1) + ql] = res[i];
  }
  __syncthreads();
  // ---- write the tile channel-major + BatchNorm partial sums (one warp per channel row)
  const int q = q0 + lane;
  for (int cl = warp; cl < 32 * CI; cl += kAggWarps) {
    const int c = c0 + cl;
    if (c >= a.C) break;
    float v = 0.f;
    if (q < a.M) {
      v = s_out[(size_t)cl * (kTile + 1) + lane];
      a.out[((size_t)b * a.C + c) * a.M + q] = v;
    }
    if (a.partial) {
      const float s1 = warp_sum(v), s2 = warp_sum(v * v);
      if (lane == 0) {
        a.partial[((size_t)blockIdx.x * 2 + 0) * a.C + c] = s1;
        a.partial[((size_t)blockIdx.x * 2 + 1) * a.C + c] = s2;
      }
    }
  }
}

// =================================================================================================
// backward (gather form over the CSR lists)
// =================================================================================================
template <int FAM>
__host__ __device__ constexpr int params_per_channel(int nkp) {
  return FAM == CL3D_FAM_ADAPTIVE_DP ? 4 : (FAM == CL3D_FAM_PSEUDOGRID ? nkp : 0);
}

template <int FAM, int CI>
__global__ void __launch_bounds__(kAggWarps * 32) agg_bwd_kernel(const AggArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * 32 * CI;
  const int chunkC = min(32 * CI, a.Cp - c0);
  const uint32_t row_bytes = (uint32_t)chunkC * 4u;
  constexpr int NACC = FAM == CL3D_FAM_PSEUDOGRID ? kMaxKP : (FAM == CL3D_FAM_ADAPTIVE_DP ? 4 : 1);
  const int ppc = params_per_channel<FAM>(a.nkp);
  const int R = a.rows_per_stage;
  const SmemLayout L = smem_layout(R, 32 * CI, FAM == CL3D_FAM_PSEUDOGRID, ppc * 32 * CI);
  float* s_stage = reinterpret_cast<float*>(smem + L.stage_off + (size_t)warp * kStageBytes);
  float4* s_dp = reinterpret_cast<float4*>(smem + L.dp_off) + (size_t)warp * R;
  int* s_q = reinterpret_cast<int*>(smem + L.idx_off) + (size_t)warp * R;
  float* s_h = reinterpret_cast<float*>(smem + L.h_off) + (size_t)warp * R * kMaxKP;
  float* s_out = reinterpret_cast<float*>(smem + L.out_off);
  float* s_red = reinterpret_cast<float*>(smem + L.red_off);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off) + warp;

  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  uint32_t phase = 0;

  LaneParams<FAM, CI> lp;
  load_lane_params<FAM, CI>(lp, a, c0, lane);
  float wk[FAM == CL3D_FAM_PSEUDOGRID ? kMaxKP : 1][CI];
  if (FAM == CL3D_FAM_PSEUDOGRID) {
#pragma unroll
    for (int kp = 0; kp < kMaxKP; ++kp)
#pragma unroll
      for (int i = 0; i < CI; ++i) {
        const int c = c0 + lane + 32 * i;
        wk[kp][i] = (kp < a.nkp && c < a.C) ? a.p1[(size_t)kp * a.C + c] : 0.f;
      }
  }
  // per-lane parameter-gradient accumulators over all points this warp handles
  float pacc[(FAM == CL3D_FAM_ADAPTIVE_DP || FAM == CL3D_FAM_PSEUDOGRID) ? NACC : 1][CI];
#pragma unroll
  for (int s = 0; s < ((FAM == CL3D_FAM_ADAPTIVE_DP || FAM == CL3D_FAM_PSEUDOGRID) ? NACC : 1); ++s)
#pragma unroll
    for (int i = 0; i < CI; ++i) pacc[s][i] = 0.f;

  const int tiles_per_cloud = (a.N + kTile - 1) / kTile;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int b = tile / tiles_per_cloud;
    const int j0 = (tile % tiles_per_cloud) * kTile;
    const float* gpm = a.g_pm + (size_t)b * a.M * a.Cp + c0;
    const float* qxyz = a.query_xyz + (size_t)b * a.M * 3;
    const int* ncnt = a.ncount + (size_t)b * a.M;
    const int* off = a.csr_off + (size_t)b * (a.N + 1);
    const int* ent = a.csr_ent + (size_t)b * a.M * a.K;

    for (int jl = warp; jl < kTile; jl += kAggWarps) {
      const int j = j0 + jl;
      float res[CI];
#pragma unroll
      for (int i = 0; i < CI; ++i) res[i] = 0.f;
      if (j < a.N) {
        const int e0 = off[j], e1 = off[j + 1];
        const float* sp = a.support_xyz + ((size_t)b * a.N + j) * 3;
        const float px = sp[0], py = sp[1], pz = sp[2];
        float acc[NACC][CI];
#pragma unroll
        for (int s = 0; s < NACC; ++s)
#pragma unroll
          for (int i = 0; i < CI; ++i) acc[s][i] = 0.f;

        for (int eb = e0; eb < e1; eb += R) {
          const int rows = min(R, e1 - eb);
          for (int r = lane; r < rows; r += 32) {
            const int code = ent[eb + r];
            const int q = code / a.K;
            s_q[r] = q;
            float dx = __fsub_rn(px, qxyz[q * 3 + 0]), dy = __fsub_rn(py, qxyz[q * 3 + 1]),
                  dz = __fsub_rn(pz, qxyz[q * 3 + 2]);
            if (a.normalize) {
              dx = __fmul_rn(dx, a.inv_radius);
              dy = __fmul_rn(dy, a.inv_radius);
              dz = __fmul_rn(dz, a.inv_radius);
            }
            const float scale = a.reduction == CL3D_REDUCE_AVG ? __fdiv_rn(1.f, (float)ncnt[q]) : 1.f;
            s_dp[r] = make_float4(dx, dy, dz, scale);
            if (FAM == CL3D_FAM_PSEUDOGRID) {
              for (int kp = 0; kp < a.nkp; ++kp)
                s_h[r * kMaxKP + kp] = pg_influence(dx, dy, dz, a.p0 + kp * 3, a.influence, a.extent);
              for (int kp = a.nkp; kp < kMaxKP; ++kp) s_h[r * kMaxKP + kp] = 0.f;
            }
          }
          if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)rows * row_bytes);
          __syncwarp();
          for (int r = lane; r < rows; r += 32)
            bulk_g2s(s_stage + (size_t)r * chunkC, gpm + (size_t)s_q[r] * a.Cp, row_bytes, bar);
          mbar_wait(bar, phase);
          phase ^= 1u;
          for (int r = 0; r < rows; ++r) {
            const float4 dp = s_dp[r];
            const float* row = s_stage + (size_t)r * chunkC;
            if (FAM == CL3D_FAM_PSEUDOGRID) {
              float v[CI];
#pragma unroll
              for (int i = 0; i < CI; ++i) v[i] = (lane + 32 * i < chunkC) ? row[lane + 32 * i] : 0.f;
              const float* hk = s_h + (size_t)r * kMaxKP;
#pragma unroll
              for (int kp = 0; kp < kMaxKP; ++kp) {
                const float h = hk[kp];
#pragma unroll
                for (int i = 0; i < CI; ++i) acc[kp][i] = fmaf(h, v[i], acc[kp][i]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < CI; ++i) {
                if (lane + 32 * i < chunkC) {
                  const float gs = row[lane + 32 * i] * dp.w;
                  if (FAM == CL3D_FAM_ADAPTIVE_DP) {  // S_x, S_y, S_z, S_1
                    acc[0][i] = fmaf(gs, dp.x, acc[0][i]);
                    acc[1][i] = fmaf(gs, dp.y, acc[1][i]);
                    acc[2][i] = fmaf(gs, dp.z, acc[2][i]);
                    acc[3][i] += gs;
                  } else {
                    acc[0][i] = fmaf(gs, family_weight<FAM, CI>(lp, i, dp), acc[0][i]);
                  }
                }
              }
            }
          }
          __syncwarp();
        }
        // ---- epilogue: gradient w.r.t. this support point's features, parameter gradients
        if (FAM == CL3D_FAM_ADAPTIVE_DP || FAM == CL3D_FAM_PSEUDOGRID) {
          const float* frow = a.feat_pm + ((size_t)b * a.N + j) * a.Cp + c0;
#pragma unroll
          for (int i = 0; i < CI; ++i) {
            const float f = (lane + 32 * i < chunkC) ? frow[lane + 32 * i] : 0.f;
            if (FAM == CL3D_FAM_ADAPTIVE_DP) {
              res[i] = fmaf(lp.c[i], acc[2][i], fmaf(lp.b[i], acc[1][i], fmaf(lp.a[i], acc[0][i], lp.d[i] * acc[3][i])));
#pragma unroll
              for (int s = 0; s < 4; ++s) pacc[s][i] = fmaf(f, acc[s][i], pacc[s][i]);
            } else {
              float r = 0.f;
#pragma unroll
              for (int kp = 0; kp < kMaxKP; ++kp) {
                r = fmaf(wk[kp][i], acc[kp][i], r);
                pacc[kp][i] = fmaf(f, acc[kp][i], pacc[kp][i]);
              }
              res[i] = r;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < CI; ++i) res[i] = acc[0][i];
        }
      }
#pragma unroll
      for (int i = 0; i < CI; ++i) s_out[(size_t)(lane + 32 * i) * (kTile + 1) + jl] = res[i];
    }
    __syncthreads();
    const int j = j0 + lane;
    for (int cl = warp; cl < 32 * CI; cl += kAggWarps) {
      const int c = c0 + cl;
      if (c >= a.C) break;
      if (j < a.N) a.out[((size_t)b * a.C + c) * a.N + j] = s_out[(size_t)cl * (kTile + 1) + lane];
    }
    __syncthreads();
  }
  // ---- CTA-level reduction of the parameter-gradient accumulators (warps take turns: fixed order)
  if (FAM == CL3D_FAM_ADAPTIVE_DP || FAM == CL3D_FAM_PSEUDOGRID) {
    for (int w = 0; w < kAggWarps; ++w) {
      if (warp == w) {
        for (int s = 0; s < ppc; ++s)
#pragma unroll
          for (int i = 0; i < CI; ++i) {
            float* p = s_red + (size_t)s * 32 * CI + lane + 32 * i;
            // pacc is indexed with a runtime s only through this unrolled select
            float v = 0.f;
#pragma unroll
            for (int ss = 0; ss < NACC; ++ss) v = (ss == s) ? pacc[ss][i] : v;
            *p = (w == 0) ? v : (*p + v);
          }
      }
      __syncthreads();
    }
    // partial layout: (gridDim.x, ppc, C)
    for (int e = threadIdx.x; e < ppc * 32 * CI; e += blockDim.x) {
      const int s = e / (32 * CI), cl = e % (32 * CI);
      const int c = c0 + cl;
      if (c < a.C) a.partial[((size_t)blockIdx.x * ppc + s) * a.C + c] = s_red[e];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
template <int FAM, int CI>
static int launch_fwd(const AggArgs& a, cudaStream_t stream) {
  const int nchunks = ceil_div(a.Cp, 32 * CI);
  const SmemLayout L = smem_layout(a.K, 32 * CI, FAM == CL3D_FAM_PSEUDOGRID, 0);
  cudaFuncSetAttribute(agg_fwd_kernel<FAM, CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  dim3 grid(a.ntiles, nchunks);
  agg_fwd_kernel<FAM, CI><<<grid, kAggWarps * 32, L.total, stream>>>(a); CL3D_LAUNCHED(1);
  return check_launch("agg_fwd_kernel");
}

template <int FAM, int CI>
static int launch_bwd(const AggArgs& a, int grid_x, cudaStream_t stream) {
  const int nchunks = ceil_div(a.Cp, 32 * CI);
  const int ppc = params_per_channel<FAM>(a.nkp);
  const SmemLayout L = smem_layout(a.rows_per_stage, 32 * CI, FAM == CL3D_FAM_PSEUDOGRID, ppc * 32 * CI);
  cudaFuncSetAttribute(agg_bwd_kernel<FAM, CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  dim3 grid(grid_x, nchunks);
  agg_bwd_kernel<FAM, CI><<<grid, kAggWarps * 32, L.total, stream>>>(a); CL3D_LAUNCHED(1);
  return check_launch("agg_bwd_kernel");
}

static int pick_ci(int family, int Cp) {
  int ci = ceil_div(Cp, 32);
  const int cap = family == CL3D_FAM_PSEUDOGRID ? 3 : kMaxCI;  // PseudoGrid keeps nkp accumulators per channel
  return ci > cap ? cap : ci;
}

#define DISPATCH_CI(FN, FAM, ...)                          \
  switch (ci) {                                            \
    case 1: return FN<FAM, 1>(__VA_ARGS__);                \
    case 2: return FN<FAM, 2>(__VA_ARGS__);                \
    case 3: return FN<FAM, 3>(__VA_ARGS__);                \
    case 4: return FN<FAM, 4>(__VA_ARGS__);                \
    case 5: return FN<FAM, 5>(__VA_ARGS__);                \
    default: return FN<FAM, 6>(__VA_ARGS__);               \
  }

static int dispatch_fwd(int family, int ci, const AggArgs& a, cudaStream_t s) {
  switch (family) {
    case CL3D_FAM_POSPOOL_XYZ: DISPATCH_CI(launch_fwd, CL3D_FAM_POSPOOL_XYZ, a, s)
    case CL3D_FAM_POSPOOL_SINCOS: DISPATCH_CI(launch_fwd, CL3D_FAM_POSPOOL_SINCOS, a, s)
    case CL3D_FAM_ADAPTIVE_DP: DISPATCH_CI(launch_fwd, CL3D_FAM_ADAPTIVE_DP, a, s)
    case CL3D_FAM_PSEUDOGRID: DISPATCH_CI(launch_fwd, CL3D_FAM_PSEUDOGRID, a, s)
  }
  return CL3D_ERR_UNSUPPORTED;
}
static int dispatch_bwd(int family, int ci, const AggArgs& a, int gx, cudaStream_t s) {
  switch (family) {
    case CL3D_FAM_POSPOOL_XYZ: DISPATCH_CI(launch_bwd, CL3D_FAM_POSPOOL_XYZ, a, gx, s)
    case CL3D_FAM_POSPOOL_SINCOS: DISPATCH_CI(launch_bwd, CL3D_FAM_POSPOOL_SINCOS, a, gx, s)
    case CL3D_FAM_ADAPTIVE_DP: DISPATCH_CI(launch_bwd, CL3D_FAM_ADAPTIVE_DP, a, gx, s)
    case CL3D_FAM_PSEUDOGRID: DISPATCH_CI(launch_bwd, CL3D_FAM_PSEUDOGRID, a, gx, s)
  }
  return CL3D_ERR_UNSUPPORTED;
}

static int bwd_grid_x(int ntiles) {
  const int cap = sm_count() * 2;
  return ntiles < cap ? (ntiles > 0 ? ntiles : 1) : cap;
}

}  // namespace cl3d

using namespace cl3d;

extern "C" int cl3d_agg_num_tiles(int B, int M) { return B * ceil_div(M, kTile); }

extern "C" int cl3d_agg_bwd_num_blocks(int B, int N) { return bwd_grid_x(B * ceil_div(N, kTile)); }

extern "C" int cl3d_agg_num_params(int family, int C, int shared, int nkp) {
  (void)shared;
  if (family == CL3D_FAM_ADAPTIVE_DP) return 4 * C;  // per-channel (x,y,z,bias); caller folds `shared` groups
  if (family == CL3D_FAM_PSEUDOGRID) return nkp * C;
  return 0;
}

static int check_common(int family, int reduction, int B, int N, int M, int K, int C, int shared, int nkp) {
  CL3D_REQUIRE(family >= 0 && family <= 3, "cl3d_agg: unknown family %d", family);
  CL3D_REQUIRE(reduction == CL3D_REDUCE_AVG || reduction == CL3D_REDUCE_SUM,
               "cl3d_agg: fused kernels implement avg / sum reductions (got %d)", reduction);
  CL3D_REQUIRE(B >= 0 && N >= 1 && M >= 1 && K >= 1 && C >= 1, "cl3d_agg: bad sizes");
  if (family == CL3D_FAM_POSPOOL_XYZ) CL3D_REQUIRE(C % 3 == 0, "PosPool xyz needs C %% 3 == 0 (got %d)", C);
  if (family == CL3D_FAM_POSPOOL_SINCOS) CL3D_REQUIRE(C % 6 == 0, "PosPool sin_cos needs C %% 6 == 0 (got %d)", C);
  if (family == CL3D_FAM_ADAPTIVE_DP) CL3D_REQUIRE(shared >= 1 && C % shared == 0, "AdaptiveWeight: bad shared_channels");
  if (family == CL3D_FAM_PSEUDOGRID) CL3D_REQUIRE(nkp >= 1 && nkp <= kMaxKP, "PseudoGrid: 1..%d kernel points", kMaxKP);
  return CL3D_OK;
}

extern "C" int cl3d_agg_fwd(int family, int reduction, const float* feat_pm, const float* query_xyz,
                            const float* support_xyz, const int* idx, const int* ncount, const float* p0,
                            const float* p1, int B, int N, int M, int K, int C, float radius, int normalize,
                            int shared, int nkp, float extent, int influence, float* agg, float* bn_partial,
                            cl3d_stream_t stream_) {
  int rc = check_common(family, reduction, B, N, M, K, C, shared, nkp);
  if (rc) return rc;
  CL3D_REQUIRE(feat_pm && query_xyz && support_xyz && idx && ncount && agg, "cl3d_agg_fwd: null pointer");
  if (B == 0) return CL3D_OK;
  AggArgs a = {};
  a.feat_pm = feat_pm;
  a.query_xyz = query_xyz;
  a.support_xyz = support_xyz;
  a.idx = idx;
  a.ncount = ncount;
  a.p0 = p0;
  a.p1 = p1;
  a.out = agg;
  a.partial = bn_partial;
  a.B = B; a.N = N; a.M = M; a.K = K; a.C = C;
  a.Cp = padded_channels(C);
  a.reduction = reduction;
  a.normalize = normalize;
  a.shared = shared > 0 ? shared : 1;
  a.nkp = nkp;
  a.influence = influence;
  a.inv_radius = 1.0f / radius;
  a.extent = extent;
  a.ntiles = B * ceil_div(M, kTile);
  const int ci = pick_ci(family, a.Cp);
  const int chunk = 32 * ci < a.Cp ? 32 * ci : a.Cp;
  a.rows_per_stage = kStageBytes / (chunk * 4);
  if (a.rows_per_stage > K) a.rows_per_stage = K;
  return dispatch_fwd(family, ci, a, (cudaStream_t)stream_);
}

extern "C" int cl3d_agg_bwd(int family, int reduction, const float* g_pm, const float* feat_pm,
                            const float* query_xyz, const float* support_xyz, const int* ncount,
                            const int* csr_off, const int* csr_ent, const float* p0, const float* p1, int B, int N,
                            int M, int K, int C, float radius, int normalize, int shared, int nkp, float extent,
                            int influence, float* grad_feat, float* grad_params_partial, cl3d_stream_t stream_) {
  int rc = check_common(family, reduction, B, N, M, K, C, shared, nkp);
  if (rc) return rc;
  CL3D_REQUIRE(g_pm && query_xyz && support_xyz && ncount && csr_off && csr_ent && grad_feat,
               "cl3d_agg_bwd: null pointer");
  const bool has_params = family == CL3D_FAM_ADAPTIVE_DP || family == CL3D_FAM_PSEUDOGRID;
  CL3D_REQUIRE(!has_params || (feat_pm && grad_params_partial), "cl3d_agg_bwd: family needs feat_pm and a partial buffer");
  if (B == 0) return CL3D_OK;
  AggArgs a = {};
  a.feat_pm = feat_pm;
  a.g_pm = g_pm;
  a.query_xyz = query_xyz;
  a.support_xyz = support_xyz;
  a.ncount = ncount;
  a.csr_off = csr_off;
  a.csr_ent = csr_ent;
  a.p0 = p0;
  a.p1 = p1;
  a.out = grad_feat;
  a.partial = grad_params_partial;
  a.B = B; a.N = N; a.M = M; a.K = K; a.C = C;
  a.Cp = padded_channels(C);
  a.reduction = reduction;
  a.normalize = normalize;
  a.shared = shared > 0 ? shared : 1;
  a.nkp = nkp;
  a.influence = influence;
  a.inv_radius = 1.0f / radius;
  a.extent = extent;
  a.ntiles = B * ceil_div(N, kTile);
  const int ci = pick_ci(family, a.Cp);
  const int chunk = 32 * ci < a.Cp ? 32 * ci : a.Cp;
  a.rows_per_stage = kStageBytes / (chunk * 4);
  if (a.rows_per_stage > 64) a.rows_per_stage = 64;
  return dispatch_bwd(family, ci, a, bwd_grid_x(a.ntiles), (cudaStream_t)stream_);
}
