// subsample.cu -- masked voxel-grid barycentre subsampling, bit-exact with the reference's
//   /root/reference/pytorch/ops/pt_custom_ops/_ext_src/src/masked_grid_subsampling_gpu.cu:11-153
// which runs ONE THREAD per cloud (<<<B,1>>>): bbox, voxel key per valid point, in-thread stable sort by key,
// sequential barycentre accumulation, LCG pseudo-shuffle of the voxel order + second stable sort, cyclic padding.
//
// Here one CTA (1024 threads) per cloud reproduces the same result in parallel:
//   keys      exactly the reference's integer key  iX + NX*iY + NX*NY*iZ  (same fp32 expressions)
//   grouping  dense table over the key range: count -> scan -> member lists; each group's members are sorted by
//             point index and summed SEQUENTIALLY in that order (fp32 addition is not associative: this is what
//             the reference's stable sort + running sum does), then divided by the float count
//   order     groups in ascending key order = ranks from the scan; the pseudo-shuffle keys are
//             seq[i] = (17*seq[i-1]+139) % 256, a full-period LCG (Hull-Dobell), so the stable sort by them has
//             the closed form  position(i) = base[seq[i % 256]] + i / 256
// The key table is bounded (kMaxCells); finer grids, and the (rounding-induced) case of a negative first key,
// take an exact serial path on one thread, like the reference.
#include "common.cuh"

namespace cl3d {

constexpr int kSubThreads = 1024;
constexpr int kMaxCells = 1 << 22;  // dense key table entries per cloud (16 MB of int32)

struct SubWs {  // per-cloud workspace slices
  int* table;    // [cells_cap]     count per key, then member-list cursor
  int* rank;     // [cells_cap]     group rank (ascending key) -- reuses scan output
  int* moff;     // [cells_cap]     member-list offset per key
  int* key;      // [n]             key per point
  int* member;   // [n]             point indices grouped by key
  int* gkey;     // [n]             key of group g
  float* bary;   // [3n]            barycentre of group g
  int* tmp;      // [4n]            serial-path scratch
};

__device__ __forceinline__ int block_scan_exclusive(int v, int* s_warp, int& total) {
  // exclusive scan of one int per thread over the CTA; returns this thread's prefix, total in `total`
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += t;
  }
  __syncthreads();
  if (lane == 31) s_warp[w] = x;
  __syncthreads();
  if (w == 0) {
    int y = lane < (blockDim.x >> 5) ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, y, o);
      if (lane >= o) y += t;
    }
    s_warp[lane] = y;
  }
  __syncthreads();
  total = s_warp[31];
  return x - v + (w > 0 ? s_warp[w - 1] : 0);
}

// exact serial restatement (one thread), used for pathological inputs only
__device__ void serial_stable_sort(int* key, int* val, int n, int* kb, int* vb) {
  for (int width = 1; width < n; width *= 2) {
    for (int lo = 0; lo < n; lo += 2 * width) {
      int mid = min(lo + width, n), hi = min(lo + 2 * width, n);
      int i = lo, j = mid, o = lo;
      while (i < mid && j < hi) {
        if (key[j] < key[i]) { kb[o] = key[j]; vb[o++] = val[j++]; }
        else                 { kb[o] = key[i]; vb[o++] = val[i++]; }
      }
      while (i < mid) { kb[o] = key[i]; vb[o++] = val[i++]; }
      while (j < hi)  { kb[o] = key[j]; vb[o++] = val[j++]; }
    }
    for (int i = 0; i < n; ++i) { key[i] = kb[i]; val[i] = vb[i]; }
  }
}

__global__ void __launch_bounds__(kSubThreads) grid_subsample_kernel(const float* __restrict__ points,
                                                                     const int* __restrict__ mask, int n, int m,
                                                                     float dl, int cells_cap, int* __restrict__ ws_i,
                                                                     float* __restrict__ sub_xyz,
                                                                     int* __restrict__ sub_mask) {
  const int b = blockIdx.x;
  const float* d = points + (size_t)b * n * 3;
  const int* mk = mask + (size_t)b * n;
  float* sx = sub_xyz + (size_t)b * m * 3;
  int* smk = sub_mask + (size_t)b * m;
  // workspace carve-up (ints): table, moff: cells_cap each; key, member, gkey: n each; bary: 3n; tmp: 4n
  const size_t per = (size_t)2 * cells_cap + (size_t)10 * n;
  int* base = ws_i + (size_t)b * per;
  int* table = base;
  int* moff = base + cells_cap;
  int* key = moff + cells_cap;
  int* member = key + n;
  int* gkey = member + n;
  float* bary = reinterpret_cast<float*>(gkey + n);
  int* tmp = reinterpret_cast<int*>(bary + 3 * (size_t)n);

  __shared__ float s_red[6][32];
  __shared__ int s_warp[32];
  __shared__ int s_nvalid, s_end, s_serial;
  __shared__ float s_org[3];
  __shared__ int s_dim[3];
  __shared__ int s_klo, s_cells;
  __shared__ int s_seq[256], s_base[257];

  // ---- n_valid (first mask 0), bbox over ALL n rows (padding included, :31-46)
  if (threadIdx.x == 0) { s_nvalid = n; s_serial = 0; }
  __syncthreads();
  {
    int first = n;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (mk[i] == 0) { first = i; break; }
    if (first < n) atomicMin(&s_nvalid, first);
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float v = d[i * 3 + a];
        mn[a] = fminf(mn[a], v);
        mx[a] = fmaxf(mx[a], v);
      }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
        mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
      }
      if ((threadIdx.x & 31) == 0) {
        s_red[a][threadIdx.x >> 5] = mn[a];
        s_red[3 + a][threadIdx.x >> 5] = mx[a];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int a = 0; a < 3; ++a)
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        s_red[a][0] = fminf(s_red[a][0], s_red[a][w]);
        s_red[3 + a][0] = fmaxf(s_red[3 + a][0], s_red[3 + a][w]);
      }
    const float inv = __fdiv_rn(1.f, dl);  // :48-50  origin = floor(min * (1/dl)) * dl
    long long cells = 1;
    for (int a = 0; a < 3; ++a) {
      s_org[a] = __fmul_rn(floorf(__fmul_rn(s_red[a][0], inv)), dl);
      s_dim[a] = (int)floorf(__fdiv_rn(__fsub_rn(s_red[3 + a][0], s_org[a]), dl)) + 1;  // :52-54
    }
    // key = iX + NX*iY + NX*NY*iZ with iX in [-1, NX], iY in [-1, NY], iZ in [-1, NZ] (rounding may push a
    // coordinate one cell outside): dense table over [klo, khi]
    const long long NX = s_dim[0], NY = s_dim[1], NZ = s_dim[2];
    const long long klo = -(1 + NX + NX * NY), khi = NX + NX * NY + NX * NY * NZ;
    cells = khi - klo + 1;
    if (NX <= 0 || NY <= 0 || NZ <= 0 || cells > cells_cap || cells <= 0) {
      s_serial = 1;
      s_klo = 0;
      s_cells = 0;
    } else {
      s_klo = (int)klo;
      s_cells = (int)cells;
    }
  }
  __syncthreads();
  const int nv = s_nvalid;
  const float ox = s_org[0], oy = s_org[1], oz = s_org[2];
  const int NX = s_dim[0], NY = s_dim[1];
  const int klo = s_klo, cells = s_cells;

  // ---- keys (:59-76), group counts
  for (int i = threadIdx.x; i < cells; i += blockDim.x) table[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const int iX = (int)floorf(__fdiv_rn(__fsub_rn(d[i * 3 + 0], ox), dl));
    const int iY = (int)floorf(__fdiv_rn(__fsub_rn(d[i * 3 + 1], oy), dl));
    const int iZ = (int)floorf(__fdiv_rn(__fsub_rn(d[i * 3 + 2], oz), dl));
    const int k = iX + NX * iY + NX * NY * iZ;
    key[i] = k;
    if (!s_serial) {
      const int t = k - klo;
      if (t < 0 || t >= cells) s_serial = 1;  // cannot happen for finite input; exact serial path if it does
      else atomicAdd(&table[t], 1);
    }
  }
  __syncthreads();

  int end = 0;
  if (nv == 0) {
    end = 0;
  } else if (s_serial) {
    // ---- exact serial path on one thread (the reference's own sequence, :77-135)
    if (threadIdx.x == 0) {
      int* kb = tmp;
      int* vb = tmp + n;
      int* k2 = tmp + 2 * n;
      int* v2 = tmp + 3 * n;
      for (int i = 0; i < nv; ++i) member[i] = i;
      serial_stable_sort(key, member, nv, kb, vb);
      int top = 0, cur = key[0], j = member[0];
      float xs = d[j * 3], ys = d[j * 3 + 1], zs = d[j * 3 + 2], pn = 1.f;
      for (int i = 1; i < nv; ++i) {
        j = member[i];
        if (key[i] == cur) {
          xs += d[j * 3]; ys += d[j * 3 + 1]; zs += d[j * 3 + 2]; pn += 1.f;
        } else {
          bary[top * 3] = __fdiv_rn(xs, pn); bary[top * 3 + 1] = __fdiv_rn(ys, pn); bary[top * 3 + 2] = __fdiv_rn(zs, pn);
          ++top;
          xs = d[j * 3]; ys = d[j * 3 + 1]; zs = d[j * 3 + 2]; pn = 1.f;
          cur = key[i];
        }
      }
      bary[top * 3] = __fdiv_rn(xs, pn); bary[top * 3 + 1] = __fdiv_rn(ys, pn); bary[top * 3 + 2] = __fdiv_rn(zs, pn);
      ++top;
      k2[0] = key[0] % 256;
      v2[0] = 0;
      for (int i = 1; i < top; ++i) { k2[i] = (17 * k2[i - 1] + 139) % 256; v2[i] = i; }
      serial_stable_sort(k2, v2, top, kb, vb);
      for (int i = 0; i < top && i < m; ++i) {
        const int g = v2[i];
        sx[i * 3] = bary[g * 3]; sx[i * 3 + 1] = bary[g * 3 + 1]; sx[i * 3 + 2] = bary[g * 3 + 2];
        smk[i] = 1;
      }
      s_end = top;
    }
    __syncthreads();
    end = s_end;
  } else {
    // ---- scan over the key table: group rank (ascending key) and member-list offsets
    const int per_t = (cells + blockDim.x - 1) / blockDim.x;
    const int lo = min((int)threadIdx.x * per_t, cells), hi = min(lo + per_t, cells);
    int occ = 0, cnt = 0;
    for (int i = lo; i < hi; ++i) {
      const int c = table[i];
      occ += c > 0;
      cnt += c;
    }
    int tot_occ, tot_cnt;
    int pocc = block_scan_exclusive(occ, s_warp, tot_occ);
    int pcnt = block_scan_exclusive(cnt, s_warp, tot_cnt);
    for (int i = lo; i < hi; ++i) {
      const int c = table[i];
      moff[i] = pcnt;
      if (c > 0) {
        gkey[pocc] = i;  // table index of group `pocc`
        ++pocc;
      }
      pcnt += c;
      table[i] = 0;  // becomes the fill cursor
    }
    end = tot_occ;
    __syncthreads();
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      const int t = key[i] - klo;
      member[moff[t] + atomicAdd(&table[t], 1)] = i;
    }
    __syncthreads();
    // ---- barycentres: members sorted by point index, summed sequentially in that order (:84-122)
    for (int g = threadIdx.x; g < end; g += blockDim.x) {
      const int t = gkey[g];
      const int o = moff[t], c = table[t];
      int* mem = member + o;
      for (int i = 1; i < c; ++i) {  // insertion sort (groups are small)
        const int v = mem[i];
        int j = i - 1;
        while (j >= 0 && mem[j] > v) { mem[j + 1] = mem[j]; --j; }
        mem[j + 1] = v;
      }
      int j = mem[0];
      float xs = d[j * 3], ys = d[j * 3 + 1], zs = d[j * 3 + 2], pn = 1.f;
      for (int i = 1; i < c; ++i) {
        j = mem[i];
        xs = __fadd_rn(xs, d[j * 3]);
        ys = __fadd_rn(ys, d[j * 3 + 1]);
        zs = __fadd_rn(zs, d[j * 3 + 2]);
        pn += 1.f;
      }
      bary[g * 3 + 0] = __fdiv_rn(xs, pn);
      bary[g * 3 + 1] = __fdiv_rn(ys, pn);
      bary[g * 3 + 2] = __fdiv_rn(zs, pn);
    }
    // ---- pseudo-shuffle (:124-135): full-period LCG -> closed-form stable order
    if (threadIdx.x == 0) {
      const int first_key = gkey[0] + klo;
      if (first_key < 0) {
        s_serial = 2;  // C's % of a negative first key leaves the LCG's 0..255 orbit: order serially below
      } else {
        int x = first_key % 256;
        for (int i = 0; i < 256; ++i) { s_seq[i] = x; x = (17 * x + 139) % 256; }
      }
    }
    __syncthreads();
    if (s_serial == 2) {
      if (threadIdx.x == 0) {
        int* k2 = tmp + 2 * n;
        int* v2 = tmp + 3 * n;
        k2[0] = (gkey[0] + klo) % 256;
        v2[0] = 0;
        for (int i = 1; i < end; ++i) { k2[i] = (17 * k2[i - 1] + 139) % 256; v2[i] = i; }
        serial_stable_sort(k2, v2, end, tmp, tmp + n);
        for (int i = 0; i < end && i < m; ++i) {
          const int g = v2[i];
          sx[i * 3] = bary[g * 3]; sx[i * 3 + 1] = bary[g * 3 + 1]; sx[i * 3 + 2] = bary[g * 3 + 2];
          smk[i] = 1;
        }
      }
      __syncthreads();
    } else {
      if (threadIdx.x < 256) {  // number of groups carrying shuffle key v, then exclusive scan over v
        const int pos = threadIdx.x;  // seq[pos] = v appears at i = pos, pos+256, ...
        const int c = end > pos ? (end - pos + 255) / 256 : 0;
        s_base[1 + s_seq[pos]] = c;
      }
      if (threadIdx.x == 0) s_base[0] = 0;
      __syncthreads();
      if (threadIdx.x == 0)
        for (int v = 0; v < 256; ++v) s_base[v + 1] += s_base[v];
      __syncthreads();
      for (int g = threadIdx.x; g < end; g += blockDim.x) {
        const int p = s_base[s_seq[g & 255]] + (g >> 8);
        if (p < m) {
          sx[p * 3 + 0] = bary[g * 3 + 0];
          sx[p * 3 + 1] = bary[g * 3 + 1];
          sx[p * 3 + 2] = bary[g * 3 + 2];
          smk[p] = 1;
        }
      }
      __syncthreads();
    }
  }
  // ---- cyclic padding with true sub points (:146-151)
  __threadfence_block();
  __syncthreads();
  for (int i = end + threadIdx.x; i < m; i += blockDim.x) {
    const int src = end > 0 ? i % end : 0;
    sx[i * 3 + 0] = end > 0 ? sx[src * 3 + 0] : 0.f;
    sx[i * 3 + 1] = end > 0 ? sx[src * 3 + 1] : 0.f;
    sx[i * 3 + 2] = end > 0 ? sx[src * 3 + 2] : 0.f;
    smk[i] = 0;
  }
}

static int cells_cap_for(int n) {
  long long c = (long long)n * 64;
  if (c < (1 << 16)) c = 1 << 16;
  if (c > kMaxCells) c = kMaxCells;
  return (int)c;
}

}  // namespace cl3d

using namespace cl3d;

extern "C" size_t cl3d_grid_subsample_workspace_bytes(int B, int n, int m) {
  (void)m;
  if (B <= 0 || n <= 0) return 256;
  return sizeof(int) * (size_t)B * ((size_t)2 * cells_cap_for(n) + (size_t)10 * n) + 256;
}

extern "C" int cl3d_grid_subsample(const float* points, const int* mask, int B, int n, int m, float sampleDl,
                                   float* sub_xyz, int* sub_mask, void* workspace, size_t workspace_bytes,
                                   cl3d_stream_t stream_) {
  CL3D_REQUIRE(points && mask && sub_xyz && sub_mask && B >= 0 && n >= 1 && m >= 1 && sampleDl > 0.f,
               "cl3d_grid_subsample: bad arguments");
  if (B == 0) return CL3D_OK;
  if (!workspace || workspace_bytes < cl3d_grid_subsample_workspace_bytes(B, n, m)) {
    set_error("cl3d_grid_subsample: workspace too small");
    return CL3D_ERR_WORKSPACE;
  }
  grid_subsample_kernel<<<B, kSubThreads, 0, (cudaStream_t)stream_>>>(points, mask, n, m, sampleDl, cells_cap_for(n),
                                                                    (int*)workspace, sub_xyz, sub_mask);
  CL3D_LAUNCHED(1);
  return check_launch("grid_subsample_kernel");
}
