// agg_common.cuh -- argument block and tile constants shared by the fused aggregation kernels (agg.cu, pg.cu).
#pragma once
#include "common.cuh"

namespace cl3d {

constexpr int kAggWarps = 8;
constexpr int kTile = 32;   // queries (fwd) / support points (bwd) per CTA tile
constexpr int kMaxKP = 16;  // PseudoGrid kernel points (reference default 15)
constexpr int kMaxCI = 6;   // channel chunk = 32*CI <= 192 channels per CTA
constexpr int kSlots = 32;  // neighbour slots (fwd) / CSR entries (bwd) staged per round

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
  int ntiles;
  unsigned char* arg_pm;      // (B,M,Cp) max reduction only: winning neighbour slot per (query, channel)
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
    if constexpr (FAM == CL3D_FAM_POSPOOL_XYZ) {
      lp.axis[i] = c % 3;  // view(B, C//3, 3, ...) : local_aggregation_operators.py:67
    } else if constexpr (FAM == CL3D_FAM_POSPOOL_SINCOS) {
      const int F = a.C / 6;  // channel = axis*2F + t ; t<F sin, t>=F cos  (:70-83)
      const int t = c % (2 * F);
      lp.axis[i] = c / (2 * F);
      lp.is_cos[i] = t >= F;
      lp.a[i] = a.p0[t % F];
    } else if constexpr (FAM == CL3D_FAM_ADAPTIVE_DP) {
      const int g = c / a.shared;  // :194-197 channel c uses weight row c // S
      lp.a[i] = a.p0[g * 3 + 0];
      lp.b[i] = a.p0[g * 3 + 1];
      lp.c[i] = a.p0[g * 3 + 2];
      lp.d[i] = a.p1[g];
    }
  }
}

// weight of channel slot i for relative position dp (float4: x,y,z,-)
template <int FAM, int CI>
__device__ __forceinline__ float family_weight(const LaneParams<FAM, CI>& lp, int i, const float4& dp) {
  if constexpr (FAM == CL3D_FAM_POSPOOL_XYZ) {
    return lp.axis[i] == 0 ? dp.x : (lp.axis[i] == 1 ? dp.y : dp.z);
  } else if constexpr (FAM == CL3D_FAM_POSPOOL_SINCOS) {
    const float p = lp.axis[i] == 0 ? dp.x : (lp.axis[i] == 1 ? dp.y : dp.z);
    const float arg = __fdiv_rn(__fmul_rn(100.f, p), lp.a[i]);  // torch.div(alpha * dp, dim_mat) :75-77
    return lp.is_cos[i] ? cosf(arg) : sinf(arg);
  } else if constexpr (FAM == CL3D_FAM_ADAPTIVE_DP) {
    return fmaf(lp.c[i], dp.z, fmaf(lp.b[i], dp.y, fmaf(lp.a[i], dp.x, lp.d[i])));  // 1x1 conv 3 -> C/S, bias
  } else {
    return 0.f;
  }
}


// pg.cu: PseudoGrid forward / backward, float4 lanes + packed fp32 FMA + sparse kernel-point walk.
// Return CL3D_ERR_UNSUPPORTED when the shape is outside their range (the caller then uses the generic kernels).
int pg2_launch_fwd(const AggArgs& a, cudaStream_t stream);
int pg2_launch_bwd(const AggArgs& a, int grid_x, cudaStream_t stream);
bool pg2_supported(const AggArgs& a);

// agg_max.cu: max reduction for PosPool (xyz | sin_cos) and AdaptiveWeight
int aggmax_launch(int family, const AggArgs& a, bool bwd, int grid_x, cudaStream_t stream);

}  // namespace cl3d
