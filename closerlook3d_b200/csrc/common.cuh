// common.cuh -- shared device/host helpers for libcl3d (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cl3d.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcl3d is written for sm_100a (Blackwell) only"
#endif

namespace cl3d {

constexpr int kWarp = 32;
constexpr int kNumSMsFallback = 148;

// ---------------------------------------------------------------------------------------------
// error reporting (thread-local message, integer return codes; never exit())
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define CL3D_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::cl3d::set_error(__VA_ARGS__);      \
      return CL3D_ERR_BAD_ARG;             \
    }                                      \
  } while (0)

// true exactly once per CUDA device (host side; for per-device one-time settings such as
// cudaFuncAttributeMaxDynamicSharedMemorySize, which is a per-device attribute)
inline bool first_call_on_device(std::atomic<unsigned long long>& seen) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  return (seen.fetch_or(bit) & bit) == 0;
}
// Opt a kernel into the full dynamic shared-memory carve-out (227 KB) ONCE per device instead of calling
// cudaFuncSetAttribute on every launch.  `seen` must be a static per kernel instantiation.
constexpr int kMaxDynSmem = 227 * 1024;
template <typename K>
inline void allow_big_smem(K kernel, std::atomic<unsigned long long>& seen) {
  if (first_call_on_device(seen)) {
    // the 227 KB limit covers static + dynamic shared memory of the block
    cudaFuncAttributes fa;
    int stat = 0;
    if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) stat = (int)fa.sharedSizeBytes;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem - stat) != cudaSuccess)
      cudaGetLastError();  // reported by the launch that follows if the kernel really needs it
  }
}
__host__ __device__ inline int padded_channels(int C) { return (C + 7) & ~7; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();
int persistent_grid_cap();  // 4 x SMs (test hook: CL3D_TEST_MAX_GRID)

// number of kernels this library has launched in this process (bench.py reports it as gpu_launches)
void count_launches(int n);
#define CL3D_LAUNCHED(n) ::cl3d::count_launches(n)

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// The reference's squared distance, bit-exact: nvcc contracts the source expression of
// masked_ordered_ball_query_gpu.cu:56-57 / masked_nearest_query_gpu.cu:47-48 to
//   t = dy*dy; t = fma(dx,dx,t); t = fma(dz,dz,t)   with d = query - support   (checked in SASS).
__device__ __forceinline__ float ref_d2(float qx, float qy, float qz, float x, float y, float z) {
  float dx = __fsub_rn(qx, x), dy = __fsub_rn(qy, y), dz = __fsub_rn(qz, z);
  float t = __fmul_rn(dy, dy);
  t = __fmaf_rn(dx, dx, t);
  t = __fmaf_rn(dz, dz, t);
  return t;
}

// base + elem_off floats as ONE instruction (IMAD.WIDE.U32): the compiler otherwise rebuilds the 64-bit address
// of a gathered row from two partial bases with four instructions (measured in the PointWiseMLP gather loops)
__device__ __forceinline__ const float* row_at(const float* base, unsigned elem_off) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(r) : "r"(elem_off), "l"(base));
  return reinterpret_cast<const float*>(r);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// ---- packed fp32x2 arithmetic (Blackwell: FFMA2 / FMUL2 / FADD2).  A register-operand FFMA issues every other
// cycle per scheduler; the packed forms carry two IEEE operations in the same slot.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {  // a - b, component-wise (exactly rounded like __fsub_rn)
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ float sqrt_approx(float x) {  // MUFU: 2^-22 relative, sqrt(0) = 0
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ u64 fma2_bcast(float h, u64 b, u64 c) {  // (h, h) * b + c ; h is a scalar-broadcast operand
  u64 d;
  asm("{\n.reg .b64 hh;\nmov.b64 hh, {%1, %1};\nfma.rn.f32x2 %0, hh, %2, %3;\n}" : "=l"(d) : "f"(h), "l"(b), "l"(c));
  return d;
}

// The same operations on plain float operands (32-bit "f" constraints).  ptxas then allocates the register pairs
// itself and keeps loop-carried accumulators in place; with 64-bit "l" operands it computed into the freshly loaded
// pair and copied the result back with four MOVs per step (profiles/r2e).
//   (cx, cy) += (h, h) * (bx, by)
__device__ __forceinline__ void ffma2_bcast(float h, float bx, float by, float& cx, float& cy) {
  asm("{\n.reg .b64 a, b, hh;\nmov.b64 a, {%0, %1};\nmov.b64 b, {%3, %4};\nmov.b64 hh, {%2, %2};\n"
      "fma.rn.f32x2 a, hh, b, a;\nmov.b64 {%0, %1}, a;\n}"
      : "+f"(cx), "+f"(cy)
      : "f"(h), "f"(bx), "f"(by));
}
//   (cx, cy) += (ax, ay) * (bx, by)
__device__ __forceinline__ void ffma2(float ax, float ay, float bx, float by, float& cx, float& cy) {
  asm("{\n.reg .b64 a, b, c;\nmov.b64 a, {%2, %3};\nmov.b64 b, {%4, %5};\nmov.b64 c, {%0, %1};\n"
      "fma.rn.f32x2 c, a, b, c;\nmov.b64 {%0, %1}, c;\n}"
      : "+f"(cx), "+f"(cy)
      : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
//   (cx, cy) = (ax, ay) * (bx, by)
__device__ __forceinline__ void fmul2(float ax, float ay, float bx, float by, float& cx, float& cy) {
  asm("{\n.reg .b64 a, b, c;\nmov.b64 a, {%2, %3};\nmov.b64 b, {%4, %5};\n"
      "mul.rn.f32x2 c, a, b;\nmov.b64 {%0, %1}, c;\n}"
      : "=f"(cx), "=f"(cy)
      : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
// ---- mbarrier helpers (producer/consumer rings of the tcgen05 GEMM, gemm_tc.cuh) ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
}  // namespace cl3d
