"""Thin tensor-level wrappers over the C ABI (include/cl3d.h): allocate outputs/workspaces with torch,
pass raw device pointers + the current CUDA stream.  No computation happens in Python."""
import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr

I32 = torch.int32
F32 = torch.float32


def padded_channels(C):
    return (C + 7) & ~7


def ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample, want_mask=True, want_ncount=True,
               algo=0):
    """-> (idx (B,M,K) i32, idx_mask (B,M,K) i32 | None, ncount (B,M) i32 | None); bit-exact with the
    reference's masked_ordered_ball_query (masked_ordered_ball_query_gpu.cu:11-96)."""
    require_cuda(query_xyz, "query_xyz", F32)
    require_cuda(support_xyz, "support_xyz", F32)
    require_cuda(query_mask, "query_mask", I32)
    require_cuda(support_mask, "support_mask", I32)
    B, M, _ = query_xyz.shape
    N = support_xyz.shape[1]
    K = int(nsample)
    dev = query_xyz.device
    L = _lib.lib()
    idx = torch.empty(B, M, K, dtype=I32, device=dev)
    idx_mask = torch.empty(B, M, K, dtype=I32, device=dev) if want_mask else None
    ncount = torch.empty(B, M, dtype=I32, device=dev) if want_ncount else None
    wsb = L.cl3d_ball_query_workspace_bytes(B, N, M, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(L.cl3d_ball_query_algo(ptr(query_xyz), ptr(support_xyz), ptr(query_mask), ptr(support_mask), B, N, M,
                                 float(radius), K, ptr(idx), ptr(idx_mask), ptr(ncount), ptr(ws), wsb, int(algo),
                                 stream_ptr()), "cl3d_ball_query")
    return idx, idx_mask, ncount


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
    check(_lib.lib().cl3d_nearest_query(ptr(query_xyz), ptr(support_xyz), ptr(query_mask), ptr(support_mask), B, N, M,
                                        ptr(idx), ptr(idx_mask), stream_ptr()), "cl3d_nearest_query")
    return idx, idx_mask


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


def to_point_major(x_cn):
    """(B,C,N) -> (B,N,Cp) zero-padded rows"""
    require_cuda(x_cn, "features", F32)
    B, C, N = x_cn.shape
    out = torch.empty(B, N, padded_channels(C), dtype=F32, device=x_cn.device)
    check(_lib.lib().cl3d_to_point_major(ptr(x_cn), B, C, N, ptr(out), stream_ptr()), "cl3d_to_point_major")
    return out


def to_channel_major(x_nc, C):
    """(B,N,Cp) -> (B,C,N)"""
    B, N, Cp = x_nc.shape
    assert Cp == padded_channels(C)
    out = torch.empty(B, C, N, dtype=F32, device=x_nc.device)
    check(_lib.lib().cl3d_to_channel_major(ptr(x_nc), B, C, N, ptr(out), stream_ptr()), "cl3d_to_channel_major")
    return out


def group_points(points, idx):
    require_cuda(points, "points", F32)
    require_cuda(idx, "idx", I32)
    B, C, N = points.shape
    _, M, K = idx.shape
    out = torch.empty(B, C, M, K, dtype=F32, device=points.device)
    check(_lib.lib().cl3d_group_points(ptr(points), ptr(idx), B, C, N, M, K, ptr(out), stream_ptr()),
          "cl3d_group_points")
    return out


def group_points_grad(grad_out, idx, n):
    require_cuda(grad_out, "grad_out", F32)
    require_cuda(idx, "idx", I32)
    B, C, M, K = grad_out.shape
    out = torch.empty(B, C, int(n), dtype=F32, device=grad_out.device)
    check(_lib.lib().cl3d_group_points_grad(ptr(grad_out), ptr(idx), B, C, int(n), M, K, ptr(out), stream_ptr()),
          "cl3d_group_points_grad")
    return out


def reduce_partials(partial):
    """(T,P) -> (P,)"""
    T, P = partial.shape
    out = torch.empty(P, dtype=F32, device=partial.device)
    check(_lib.lib().cl3d_reduce_partials(ptr(partial), T, P, ptr(out), stream_ptr()), "cl3d_reduce_partials")
    return out
