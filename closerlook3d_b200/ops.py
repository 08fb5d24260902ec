"""Thin tensor-level wrappers over the C ABI (include/cl3d.h): allocate outputs/workspaces with torch,
pass raw device pointers + the current CUDA stream.  No computation happens in Python."""
import ctypes

import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr

I32 = torch.int32
F32 = torch.float32


def _on_device(fn):
    """Run `fn` with the CUDA device of its tensor arguments current (kernels, the stream passed to the library and
    the per-device attributes all belong to the CURRENT device), and refuse operands spread over several devices
    (the reference's ops have the same single-device contract)."""
    import functools

    @functools.wraps(fn)
    def call(*args, **kw):
        dev = None
        for a in list(args) + list(kw.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise RuntimeError(f"{fn.__name__}: operands on different devices ({dev} and {a.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            return fn(*args, **kw)
    return call


def padded_channels(C):
    return (C + 7) & ~7


PM_SLACK = 256  # CL3D_PM_SLACK: readable floats after the last row of every gathered point-major buffer


def _empty_pm(B, N, W, device):
    """(B,N,W) fp32 view of an allocation with PM_SLACK floats of slack behind it"""
    return torch.empty(B * N * W + PM_SLACK, dtype=F32, device=device)[:B * N * W].view(B, N, W)


@_on_device
def ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample, want_mask=True, want_ncount=True,
               algo=0, csr=None, after_search=None, out=None):
    """-> (idx (B,M,K) i32, idx_mask (B,M,K) i32 | None, ncount (B,M) i32 | None); bit-exact with the
    reference's masked_ordered_ball_query (masked_ordered_ball_query_gpu.cu:11-96).
    csr = "counted" | "all": also build the transposed lists in the same call (cl3d_ball_query_csr) and return
    (idx, idx_mask, ncount, (csr_off, csr_ent)).  after_search(): called between the search and the list build (the
    caller records its "search done" event there, so consumers of idx do not wait for the lists).
    out: dict of preallocated result tensors (idx, idx_mask, ncount, off, ent -- e.g. batch slices of larger tensors)
    to write into instead of allocating."""
    require_cuda(query_xyz, "query_xyz", F32)
    require_cuda(support_xyz, "support_xyz", F32)
    require_cuda(query_mask, "query_mask", I32)
    require_cuda(support_mask, "support_mask", I32)
    B, M, _ = query_xyz.shape
    N = support_xyz.shape[1]
    K = int(nsample)
    dev = query_xyz.device
    L = _lib.lib()
    out = out or {}
    idx = out["idx"] if "idx" in out else torch.empty(B, M, K, dtype=I32, device=dev)
    idx_mask = out.get("idx_mask") if "idx" in out else (torch.empty(B, M, K, dtype=I32, device=dev) if want_mask else None)
    ncount = out["ncount"] if "ncount" in out else \
        (torch.empty(B, M, dtype=I32, device=dev) if (want_ncount or csr) else None)
    if csr:
        assert csr in ("counted", "all")
        off = out["off"] if "off" in out else torch.empty(B, N + 1, dtype=I32, device=dev)
        ent = out["ent"] if "ent" in out else torch.empty(B, M * K, dtype=I32, device=dev)
        wsb = L.cl3d_ball_query_csr_workspace_bytes(B, N, M, K)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        args = (ptr(query_xyz), ptr(support_xyz), ptr(query_mask), ptr(support_mask), B, N, M, float(radius), K, ptr(idx),
                ptr(idx_mask), ptr(ncount), 1 if csr == "all" else 0, ptr(off), ptr(ent), ptr(ws), wsb, int(algo))
        if after_search is None:
            check(L.cl3d_ball_query_csr(*args, 3, stream_ptr()), "cl3d_ball_query_csr")
        else:
            check(L.cl3d_ball_query_csr(*args, 1, stream_ptr()), "cl3d_ball_query_csr (search)")
            after_search()
            check(L.cl3d_ball_query_csr(*args, 2, stream_ptr()), "cl3d_ball_query_csr (lists)")
        return idx, idx_mask, ncount, (off, ent)
    wsb = L.cl3d_ball_query_workspace_bytes(B, N, M, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(L.cl3d_ball_query_algo(ptr(query_xyz), ptr(support_xyz), ptr(query_mask), ptr(support_mask), B, N, M,
                                 float(radius), K, ptr(idx), ptr(idx_mask), ptr(ncount), ptr(ws), wsb, int(algo),
                                 stream_ptr()), "cl3d_ball_query")
    return idx, idx_mask, ncount


@_on_device
def nearest_query(query_xyz, support_xyz, query_mask, support_mask):
    """-> (idx (B,M) i32, idx_mask (B,M) i32)  (masked_nearest_query_gpu.cu:8-62)"""
    require_cuda(query_xyz, "query_xyz", F32)
    require_cuda(support_xyz, "support_xyz", F32)
    require_cuda(query_mask, "query_mask", I32)
    require_cuda(support_mask, "support_mask", I32)
    B, M, _ = query_xyz.shape
    N = support_xyz.shape[1]
    idx = torch.empty(B, M, dtype=I32, device=query_xyz.device)
    idx_mask = torch.empty(B, M, dtype=I32, device=query_xyz.device)
    L = _lib.lib()
    wsb = L.cl3d_nearest_query_workspace_bytes(B, N, M)      # 0: small cloud, tile scan; else the cell grid
    ws = torch.empty(wsb, dtype=torch.uint8, device=query_xyz.device) if wsb else None
    check(L.cl3d_nearest_query(ptr(query_xyz), ptr(support_xyz), ptr(query_mask), ptr(support_mask), B, N, M,
                               ptr(idx), ptr(idx_mask), ptr(ws), wsb, stream_ptr()), "cl3d_nearest_query")
    return idx, idx_mask


@_on_device
def build_csr(idx, ncount, N):
    """transposed neighbour lists -> (csr_off (B,N+1) i32, csr_ent (B,M*K) i32)"""
    B, M, K = idx.shape
    dev = idx.device
    L = _lib.lib()
    off = torch.empty(B, N + 1, dtype=I32, device=dev)
    ent = torch.empty(B, M * K, dtype=I32, device=dev)
    wsb = L.cl3d_csr_workspace_bytes(B, N, M, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(L.cl3d_build_csr(ptr(idx), ptr(ncount), B, N, M, K, ptr(off), ptr(ent), ptr(ws), wsb, stream_ptr()),
          "cl3d_build_csr")
    return off, ent


@_on_device
def to_point_major(x_cn):
    """(B,C,N) -> (B,N,Cp) zero-padded rows"""
    require_cuda(x_cn, "features", F32)
    B, C, N = x_cn.shape
    out = _empty_pm(B, N, padded_channels(C), x_cn.device)
    check(_lib.lib().cl3d_to_point_major(ptr(x_cn), B, C, N, ptr(out), stream_ptr()), "cl3d_to_point_major")
    return out


@_on_device
def to_channel_major(x_nc, C):
    """(B,N,Cp) -> (B,C,N)"""
    B, N, Cp = x_nc.shape
    assert Cp == padded_channels(C)
    out = torch.empty(B, C, N, dtype=F32, device=x_nc.device)
    check(_lib.lib().cl3d_to_channel_major(ptr(x_nc), B, C, N, ptr(out), stream_ptr()), "cl3d_to_channel_major")
    return out


@_on_device
def group_points(points, idx):
    require_cuda(points, "points", F32)
    require_cuda(idx, "idx", I32)
    B, C, N = points.shape
    _, M, K = idx.shape
    out = torch.empty(B, C, M, K, dtype=F32, device=points.device)
    check(_lib.lib().cl3d_group_points(ptr(points), ptr(idx), B, C, N, M, K, ptr(out), stream_ptr()),
          "cl3d_group_points")
    return out


@_on_device
def group_points_grad(grad_out, idx, n):
    require_cuda(grad_out, "grad_out", F32)
    require_cuda(idx, "idx", I32)
    B, C, M, K = grad_out.shape
    out = torch.empty(B, C, int(n), dtype=F32, device=grad_out.device)
    check(_lib.lib().cl3d_group_points_grad(ptr(grad_out), ptr(idx), B, C, int(n), M, K, ptr(out), stream_ptr()),
          "cl3d_group_points_grad")
    return out


@_on_device
def reduce_partials(partial):
    """(T,P) -> (P,)"""
    T, P = partial.shape
    out = torch.empty(P, dtype=F32, device=partial.device)
    check(_lib.lib().cl3d_reduce_partials(ptr(partial), T, P, ptr(out), stream_ptr()), "cl3d_reduce_partials")
    return out


# --------------------------------------------------------------------------------------------------
# fused aggregation + out_transform
# --------------------------------------------------------------------------------------------------
FAM_POSPOOL_XYZ, FAM_POSPOOL_SINCOS, FAM_ADAPTIVE_DP, FAM_PSEUDOGRID = 0, 1, 2, 3
REDUCE = {"avg": 0, "mean": 0, "sum": 1, "max": 2}


@_on_device
def agg_fwd(family, reduction, feat_pm, query_xyz, support_xyz, idx, ncount, p0, p1, C, radius, normalize, shared=1,
            nkp=0, extent=1.0, influence=0, want_bn_partial=True, out=None, arg=None):
    """-> (agg (B,C,M), bn_partial (ntiles,2,C) | None); out = (agg, partial) preallocated (e.g. batch slices);
    arg = (B,M,Cp) uint8 buffer for the winning slots, required by (and only by) the max reduction"""
    B, N, Cp = feat_pm.shape
    M, K = idx.shape[1], idx.shape[2]
    L = _lib.lib()
    dev = feat_pm.device
    if out is not None:
        agg, partial = out
    else:
        agg = torch.empty(B, C, M, dtype=F32, device=dev)
        partial = torch.empty(L.cl3d_agg_num_tiles(B, M), 2, C, dtype=F32, device=dev) if want_bn_partial else None
    check(L.cl3d_agg_fwd(family, reduction, ptr(feat_pm), ptr(query_xyz), ptr(support_xyz), ptr(idx), ptr(ncount),
                         ptr(p0), ptr(p1), B, N, M, K, C, float(radius), int(normalize), int(shared), int(nkp),
                         float(extent), int(influence), ptr(agg), ptr(partial), ptr(arg), stream_ptr()), "cl3d_agg_fwd")
    return agg, partial


@_on_device
def agg_bwd(family, reduction, g_pm, feat_pm, query_xyz, support_xyz, ncount, csr_off, csr_ent, p0, p1, C, N, K,
            radius, normalize, shared=1, nkp=0, extent=1.0, influence=0, arg=None):
    """-> (grad_feat (B,C,N), param_grad (P//C, C) | None)   P = 4C (adaptive: x,y,z,bias) or nkp*C;
    arg = the forward's winning slots (max reduction)"""
    B, M, Cp = g_pm.shape
    L = _lib.lib()
    dev = g_pm.device
    grad_feat = torch.empty(B, C, N, dtype=F32, device=dev)
    P = L.cl3d_agg_num_params(family, C, shared, nkp)
    partial = None
    if P > 0:
        partial = torch.empty(L.cl3d_agg_bwd_num_blocks(B, N), P, dtype=F32, device=dev)
    check(L.cl3d_agg_bwd(family, reduction, ptr(g_pm), ptr(feat_pm), ptr(query_xyz), ptr(support_xyz), ptr(ncount),
                         ptr(csr_off), ptr(csr_ent), ptr(p0), ptr(p1), B, N, M, K, C, float(radius), int(normalize),
                         int(shared), int(nkp), float(extent), int(influence), ptr(grad_feat), ptr(partial),
                         ptr(arg), stream_ptr()), "cl3d_agg_bwd")
    pg = reduce_partials(partial).view(P // C, C) if P > 0 else None
    return grad_feat, pg


@_on_device
def bn_finalize(bn_partial, C, count, eps, momentum, training, running_mean, running_var):
    """-> save_stats (2,C): mean, invstd.  Updates running stats in place when training."""
    dev = running_mean.device if running_mean is not None else bn_partial.device
    stats = torch.empty(2, C, dtype=F32, device=dev)
    nt = bn_partial.shape[0] if bn_partial is not None else 0
    check(_lib.lib().cl3d_bn_finalize(ptr(bn_partial), nt, C, int(count), float(eps), float(momentum), int(training),
                                      ptr(running_mean), ptr(running_var), ptr(stats), stream_ptr()),
          "cl3d_bn_finalize")
    return stats


@_on_device
def bn_relu_fwd(x, stats, gamma, beta):
    B, C, M = x.shape
    y = torch.empty_like(x)
    check(_lib.lib().cl3d_bn_relu_fwd(ptr(x), ptr(stats), ptr(gamma), ptr(beta), B, C, M, ptr(y), stream_ptr()),
          "cl3d_bn_relu_fwd")
    return y


@_on_device
def bn_relu_bwd(grad_y, x, stats, gamma, beta, training):
    """-> (g_pm (B,M,Cp) point-major d/d(agg), dgamma (C,), dbeta (C,))"""
    B, C, M = x.shape
    L = _lib.lib()
    dev = x.device
    partial = torch.empty(L.cl3d_agg_num_tiles(B, M), 2, C, dtype=F32, device=dev)
    dgb = torch.empty(2, C, dtype=F32, device=dev)
    g_pm = _empty_pm(B, M, padded_channels(C), dev)
    check(L.cl3d_bn_relu_bwd(ptr(grad_y), ptr(x), ptr(stats), ptr(gamma), ptr(beta), B, C, M, int(training),
                             ptr(partial), ptr(dgb), ptr(g_pm), stream_ptr()), "cl3d_bn_relu_bwd")
    return g_pm, dgb[0], dgb[1]


# --------------------------------------------------------------------------------------------------
# fp32 GEMM + fused PointWiseMLP
# --------------------------------------------------------------------------------------------------
GEMM_AUTO, GEMM_FFMA, GEMM_TC3X = 0, 1, 2


@_on_device
def sgemm(a, sa_m, sa_k, b, sb_k, sb_n, M, N, K, out=None, ldc=None, splitk=1, algo=GEMM_AUTO):
    """out[m][n] = sum_k a[m*sa_m + k*sa_k] * b[k*sb_k + n*sb_n]  (element strides; fp32 accuracy: 3xTF32 on
    the tcgen05 tensor cores, or the fp32 FMA kernel -- see cl3d.h)"""
    L = _lib.lib()
    dev = a.device
    ldc = N if ldc is None else ldc
    if out is None:
        out = _empty_pm(1, M, ldc, dev).view(M, ldc)
    wsb = L.cl3d_sgemm_workspace_bytes(M, N, splitk)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev) if splitk > 1 else None
    check(L.cl3d_sgemm_algo(ptr(a), sa_m, sa_k, ptr(b), sb_k, sb_n, M, N, K, ptr(out), ldc, splitk, ptr(ws), wsb,
                            algo, stream_ptr()), "cl3d_sgemm")
    return out


@_on_device
def to_point_major_aug(x_cn, xyz, radius):
    """(B,C,N) features + (B,N,3) xyz -> (B,N,Cpa) rows [f | xyz/r | 0], Cpa = padded(C+3)"""
    require_cuda(x_cn, "features", F32)
    B, C, N = x_cn.shape
    out = torch.empty(B, N, padded_channels(C + 3), dtype=F32, device=x_cn.device)
    check(_lib.lib().cl3d_to_point_major_aug(ptr(x_cn), ptr(xyz), B, C, N, float(radius), ptr(out), stream_ptr()),
          "cl3d_to_point_major_aug")
    return out


@_on_device
def pwmlp_prep_weights(conv_weight, gamma, C, Cout):
    """-> wcat (2Cop, Cpa) rows zero-padded to Cpa = padded(C+3), wp (Cout,3), sgn (Cout)"""
    dev = conv_weight.device
    Cop = padded_channels(Cout)
    wcat = torch.empty(2 * Cop, padded_channels(C + 3), dtype=F32, device=dev)
    wp = torch.empty(Cout, 3, dtype=F32, device=dev)
    sgn = torch.empty(Cout, dtype=F32, device=dev)
    check(_lib.lib().cl3d_pwmlp_prep_weights(ptr(conv_weight), ptr(gamma), C, Cout, ptr(wcat), ptr(wp), ptr(sgn),
                                             stream_ptr()), "cl3d_pwmlp_prep_weights")
    return wcat, wp, sgn


@_on_device
def pwmlp_weight_grad(gwcat, grad_wp, sgn, C, Cout):
    gW = torch.empty(Cout, 3 + 2 * C, 1, 1, dtype=F32, device=gwcat.device)
    check(_lib.lib().cl3d_pwmlp_weight_grad(ptr(gwcat), ptr(grad_wp), ptr(sgn), C, Cout, ptr(gW), stream_ptr()),
          "cl3d_pwmlp_weight_grad")
    return gW


@_on_device
def pwmlp_fwd_stats(ab_pm, wp, sgn, query_xyz, support_xyz, idx, Cout, radius):
    B, N, _ = ab_pm.shape
    M, K = idx.shape[1], idx.shape[2]
    L = _lib.lib()
    dev = ab_pm.device
    Cop = padded_channels(Cout)
    ysel = torch.empty(B, Cout, M, dtype=F32, device=dev)
    aq = _empty_pm(B, M, Cop, dev)
    sq = torch.empty(B, M, Cop, dtype=F32, device=dev)
    karg = torch.empty(B, M, Cop, dtype=torch.uint8, device=dev)
    partial = torch.empty(L.cl3d_agg_num_tiles(B, M), 2, Cout, dtype=F32, device=dev)
    check(L.cl3d_pwmlp_fwd_stats(ptr(ab_pm), ptr(wp), ptr(sgn), ptr(query_xyz), ptr(support_xyz), ptr(idx), B, N, M, K, Cout,
                                 float(radius), ptr(ysel), ptr(aq), ptr(sq), ptr(karg), ptr(partial), stream_ptr()),
          "cl3d_pwmlp_fwd_stats")
    return ysel, aq, sq, karg, partial


@_on_device
def pwmlp_fwd_out(ysel, stats, gamma, beta):
    B, Cout, M = ysel.shape
    out = torch.empty_like(ysel)
    check(_lib.lib().cl3d_pwmlp_fwd_out(ptr(ysel), ptr(stats), ptr(gamma), ptr(beta), B, M, Cout, ptr(out),
                                        stream_ptr()), "cl3d_pwmlp_fwd_out")
    return out


@_on_device
def pwmlp_bwd(grad_out, out, ab_pm, wp, sgn, query_xyz, support_xyz, idx, csr_off, csr_ent, ysel, aq, sq, karg, stats, gamma,
              radius, side_stream=None, training=True):
    """-> (grad_ab_pm (B,N,2Cop), grad_wp (3,Cout), dgamma (Cout), dbeta (Cout)); side_stream: a torch stream the
    library may fork onto (zero-fill + query pass run beside the gather pass), joined before it returns"""
    B, N, C2 = ab_pm.shape
    Cout, M, K = out.shape[1], out.shape[2], idx.shape[2]
    L = _lib.lib()
    dev = out.device
    partial = _empty_pm(1, 1, L.cl3d_pwmlp_bwd_scratch_floats(B, N, M, Cout), dev)  # partial sums + sc*dz rows
    dgb = torch.empty(2, Cout, dtype=F32, device=dev)
    grad_ab = torch.empty(B, N, C2, dtype=F32, device=dev)
    grad_wp = torch.empty(3, Cout, dtype=F32, device=dev)
    check(L.cl3d_pwmlp_bwd(ptr(grad_out), ptr(out), ptr(ab_pm), ptr(wp), ptr(sgn), ptr(query_xyz), ptr(support_xyz), ptr(idx),
                           ptr(csr_off), ptr(csr_ent), ptr(ysel), ptr(aq), ptr(sq), ptr(karg), ptr(stats), ptr(gamma),
                           B, N, M, K, Cout, float(radius), int(bool(training)), ptr(partial), ptr(dgb), ptr(grad_ab),
                           ptr(grad_wp),
                           stream_ptr(), ctypes.c_void_p(side_stream.cuda_stream) if side_stream is not None else None),
          "cl3d_pwmlp_bwd")
    return grad_ab, grad_wp, dgb[0], dgb[1]


# --------------------------------------------------------------------------------------------------
# grid subsampling, fused max-pool
# --------------------------------------------------------------------------------------------------
@_on_device
def grid_subsample(points, mask, npoint, sampleDl):
    """-> (sub_xyz (B,m,3) f32, sub_mask (B,m) i32); bit-exact with masked_grid_subsampling_gpu.cu:11-153"""
    require_cuda(points, "points", F32)
    require_cuda(mask, "mask", I32)
    B, n, _ = points.shape
    m = int(npoint)
    L = _lib.lib()
    dev = points.device
    sub = torch.empty(B, m, 3, dtype=F32, device=dev)
    sm = torch.empty(B, m, dtype=I32, device=dev)
    wsb = L.cl3d_grid_subsample_workspace_bytes(B, n, m)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(L.cl3d_grid_subsample(ptr(points), ptr(mask), B, n, m, float(sampleDl), ptr(sub), ptr(sm), ptr(ws), wsb,
                                stream_ptr()), "cl3d_grid_subsample")
    return sub, sm


@_on_device
def gather_max(features, idx):
    """out[b,c,q] = max_k features[b,c,idx[b,q,k]] -> (out (B,C,M), arg (B,M,Cp) uint8)"""
    require_cuda(features, "features", F32)
    require_cuda(idx, "idx", I32)
    B, C, N = features.shape
    M, K = idx.shape[1], idx.shape[2]
    feat_pm = to_point_major(features)
    out = torch.empty(B, C, M, dtype=F32, device=features.device)
    arg = torch.empty(B, M, feat_pm.shape[2], dtype=torch.uint8, device=features.device)
    check(_lib.lib().cl3d_gather_max_fwd(ptr(feat_pm), ptr(idx), B, N, M, K, C, ptr(out), ptr(arg), stream_ptr()),
          "cl3d_gather_max_fwd")
    return out, arg


@_on_device
def gather_max_grad(grad_out, idx, arg, N):
    B, C, M = grad_out.shape
    K = idx.shape[2]
    g_pm = torch.empty(B, N, padded_channels(C), dtype=F32, device=grad_out.device)
    check(_lib.lib().cl3d_gather_max_bwd(ptr(grad_out), ptr(idx), ptr(arg), B, N, M, K, C, ptr(g_pm), stream_ptr()),
          "cl3d_gather_max_bwd")
    return to_channel_major(g_pm, C)
