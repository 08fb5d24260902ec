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
};


// pg.cu: PseudoGrid forward / backward, float4 lanes + packed fp32 FMA + sparse kernel-point walk.
// Return CL3D_ERR_UNSUPPORTED when the shape is outside their range (the caller then uses the generic kernels).
int pg2_launch_fwd(const AggArgs& a, cudaStream_t stream);
int pg2_launch_bwd(const AggArgs& a, int grid_x, cudaStream_t stream);
bool pg2_supported(const AggArgs& a);

}  // namespace cl3d
