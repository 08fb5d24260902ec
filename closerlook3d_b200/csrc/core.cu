// core.cu -- error reporting, device queries, layout transposes, compatibility gather/scatter.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace cl3d {

static thread_local char g_err[512] = "";
static long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (long long)n, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
    return CL3D_ERR_LAUNCH;
  }
  return CL3D_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMsFallback;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMsFallback;
    cached[dev] = n;
  }
  return cached[dev];
}

// Grid size of the persistent (tile-loop) kernels: 4 CTAs per SM.  CL3D_TEST_MAX_GRID (read on every call, tests
// only) caps it so that small inputs run several tiles per CTA.
int persistent_grid_cap() {
  int cap = sm_count() * 4;
  const char* e = getenv("CL3D_TEST_MAX_GRID");
  if (e && *e) {
    const int v = atoi(e);
    if (v >= 1 && v < cap) cap = v;
  }
  return cap;
}

// ---------------------------------------------------------------------------------------------
// (B,C,N) channel-major -> (B,N,Cp) point-major, zero padded.  32x32 smem tile transpose; reads are
// coalesced along N, writes along C.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) to_point_major_kernel(const float* __restrict__ in, int C, int N, int Cp,
                                                             float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  in += (size_t)b * C * N;
  out += (size_t)b * N * Cp;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, n = n0 + tx;
    tile[r][tx] = (c < C && n < N) ? in[(size_t)c * N + n] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, c = c0 + tx;
    if (n < N && c < Cp) out[(size_t)n * Cp + c] = tile[tx][r];
  }
}

__global__ void __launch_bounds__(256) to_channel_major_kernel(const float* __restrict__ in, int C, int N, int Cp,
                                                               float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  in += (size_t)b * N * Cp;
  out += (size_t)b * C * N;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, c = c0 + tx;
    tile[r][tx] = (n < N && c < C) ? in[(size_t)n * Cp + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, n = n0 + tx;
    if (c < C && n < N) out[(size_t)c * N + n] = tile[tx][r];
  }
}

// ---------------------------------------------------------------------------------------------
// compatibility path: materialising gather (reference group_points_gpu.cu:13-33) and its gradient
// (:48-69).  One thread per (c, j) row of K outputs; the gradient is computed in gather form over a
// per-call scan of idx ... kept simple: these are NOT on the fused hot path.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) group_points_kernel(const float* __restrict__ points,
                                                           const int* __restrict__ idx, int C, int N, int M, int K,
                                                           float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)M * K) return;
  const float* p = points + ((size_t)b * C + c) * N;
  out[((size_t)b * C + c) * M * K + e] = p[idx[(size_t)b * M * K + e]];
}

// Deterministic scatter-add: one warp owns one (b, c) row of grad_points and walks all M*K entries in
// order, accumulating through shared-memory bins would need N floats; instead use fp32 atomics like the
// reference but on a zero-filled output.  (Order non-determinism matches the reference's own atomicAdd.)
__global__ void __launch_bounds__(256) group_points_grad_kernel(const float* __restrict__ grad_out,
                                                                const int* __restrict__ idx, int C, int N, int M,
                                                                int K, float* __restrict__ grad_points) {
  const int b = blockIdx.z, c = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)M * K) return;
  float* gp = grad_points + ((size_t)b * C + c) * N;
  atomicAdd(gp + idx[(size_t)b * M * K + e], grad_out[((size_t)b * C + c) * M * K + e]);
}

// out[p] = sum_t partial[t][p].  32 parameters x 32 tile-lanes per CTA; every tile-lane sums a fixed
// strided subset in double, then the 32 lanes are combined in a fixed order -> deterministic.
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partial, int ntiles, int P,
                                                               float* __restrict__ out) {
  __shared__ double s_acc[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + tx;
  double acc = 0.0;
  if (p < P) {
    double b = 0.0, c = 0.0, d = 0.0;
    int t = ty;
    for (; t + 96 < ntiles; t += 128) {  // four independent load streams per thread
      acc += (double)partial[(size_t)t * P + p];
      b += (double)partial[(size_t)(t + 32) * P + p];
      c += (double)partial[(size_t)(t + 64) * P + p];
      d += (double)partial[(size_t)(t + 96) * P + p];
    }
    for (; t < ntiles; t += 32) acc += (double)partial[(size_t)t * P + p];
    acc = (acc + b) + (c + d);
  }
  s_acc[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && p < P) {
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) a += s_acc[i][tx];
    out[p] = (float)a;
  }
}

}  // namespace cl3d

using namespace cl3d;

extern "C" int cl3d_version(void) { return 100; }
extern "C" const char* cl3d_last_error(void) { return g_err; }
extern "C" int cl3d_padded_channels(int C) { return padded_channels(C); }
extern "C" int cl3d_sm_count(void) { return sm_count(); }
extern "C" long long cl3d_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int cl3d_to_point_major(const float* in_cn, int B, int C, int N, float* out_nc, cl3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CL3D_REQUIRE(B >= 0 && C >= 1 && N >= 1 && in_cn && out_nc, "cl3d_to_point_major: bad arguments");
  if (B == 0) return CL3D_OK;
  const int Cp = padded_channels(C);
  dim3 grid(ceil_div(N, 32), ceil_div(Cp, 32), B);
  to_point_major_kernel<<<grid, 256, 0, stream>>>(in_cn, C, N, Cp, out_nc); CL3D_LAUNCHED(1);
  return check_launch("to_point_major_kernel");
}

extern "C" int cl3d_to_channel_major(const float* in_nc, int B, int C, int N, float* out_cn, cl3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CL3D_REQUIRE(B >= 0 && C >= 1 && N >= 1 && in_nc && out_cn, "cl3d_to_channel_major: bad arguments");
  if (B == 0) return CL3D_OK;
  const int Cp = padded_channels(C);
  dim3 grid(ceil_div(N, 32), ceil_div(C, 32), B);
  to_channel_major_kernel<<<grid, 256, 0, stream>>>(in_nc, C, N, Cp, out_cn); CL3D_LAUNCHED(1);
  return check_launch("to_channel_major_kernel");
}

extern "C" int cl3d_group_points(const float* points, const int* idx, int B, int C, int N, int M, int K, float* out,
                                 cl3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CL3D_REQUIRE(B >= 0 && C >= 1 && N >= 1 && M >= 0 && K >= 1 && points && idx && out, "cl3d_group_points: bad arguments");
  CL3D_REQUIRE(C <= 65535 && B <= 65535, "cl3d_group_points: C or B too large");
  if (B == 0 || M == 0) return CL3D_OK;
  const long long ents = (long long)M * K;
  dim3 grid((unsigned)((ents + 255) / 256), C, B);
  group_points_kernel<<<grid, 256, 0, stream>>>(points, idx, C, N, M, K, out); CL3D_LAUNCHED(1);
  return check_launch("group_points_kernel");
}

extern "C" int cl3d_group_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int M, int K,
                                      float* grad_points, cl3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CL3D_REQUIRE(B >= 0 && C >= 1 && N >= 1 && M >= 0 && K >= 1 && grad_out && idx && grad_points,
               "cl3d_group_points_grad: bad arguments");
  CL3D_REQUIRE(C <= 65535 && B <= 65535, "cl3d_group_points_grad: C or B too large");
  if (B == 0) return CL3D_OK;
  cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * C * N, stream);
  if (M == 0) return CL3D_OK;
  const long long ents = (long long)M * K;
  dim3 grid((unsigned)((ents + 255) / 256), C, B);
  group_points_grad_kernel<<<grid, 256, 0, stream>>>(grad_out, idx, C, N, M, K, grad_points); CL3D_LAUNCHED(1);
  return check_launch("group_points_grad_kernel");
}

extern "C" int cl3d_reduce_partials(const float* partial, int ntiles, int P, float* out, cl3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CL3D_REQUIRE(ntiles >= 0 && P >= 1 && partial && out, "cl3d_reduce_partials: bad arguments");
  reduce_partials_kernel<<<ceil_div(P, 32), 1024, 0, stream>>>(partial, ntiles, P, out); CL3D_LAUNCHED(1);
  return check_launch("reduce_partials_kernel");
}
