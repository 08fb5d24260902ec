"""oracle/ext.py -- the C restatement (cl3d_oracle.c) behind the reference's pybind surface.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Mirrors `pt_custom_ops._ext` (reference pytorch/ops/pt_custom_ops/_ext_src/src/bindings.cpp:6-15): same five
function names, argument order, dtypes/contiguity checks (utils.h:9-30) and return shapes, but on CPU
tensors.  `oracle.ref_loader` seeds it into sys.modules['pt_custom_ops._ext'] so the reference's unmodified
python modules run on CPU.
"""
import ctypes

import torch

from . import lib

_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)


def _fp(t):
    return ctypes.cast(t.data_ptr(), _F)


def _ip(t):
    return ctypes.cast(t.data_ptr(), _I)


def _chk(t, name, dtype):
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be a{'n int' if dtype == torch.int32 else ' float'} tensor")
    if t.device.type != "cpu":
        raise RuntimeError(f"{name}: the oracle runs on CPU tensors")


def num_threads():
    return lib().cl3d_oracle_num_threads()


def set_threads(n):
    lib().cl3d_oracle_set_threads(int(n))


def masked_ordered_ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample):
    """masked_ordered_ball_query.cpp:13-59 -> [idx (B,M,K) i32, idx_mask (B,M,K) i32]"""
    _chk(query_xyz, "query_xyz", torch.float32)
    _chk(support_xyz, "support_xyz", torch.float32)
    _chk(query_mask, "query_mask", torch.int32)
    _chk(support_mask, "support_mask", torch.int32)
    B, M, _ = query_xyz.shape
    N = support_xyz.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=torch.int32)
    idx_mask = torch.zeros(B, M, nsample, dtype=torch.int32)
    lib().cl3d_oracle_ball_query(_fp(query_xyz), _fp(support_xyz), _ip(query_mask), _ip(support_mask),
                                 B, N, M, ctypes.c_float(radius), int(nsample), _ip(idx), _ip(idx_mask))
    return [idx, idx_mask]


def masked_nearest_query(query_xyz, support_xyz, query_mask, support_mask):
    """masked_nearest_query.cpp -> [idx (B,M,1) i32, idx_mask (B,M,1) i32]"""
    _chk(query_xyz, "query_xyz", torch.float32)
    _chk(support_xyz, "support_xyz", torch.float32)
    _chk(query_mask, "query_mask", torch.int32)
    _chk(support_mask, "support_mask", torch.int32)
    B, M, _ = query_xyz.shape
    N = support_xyz.shape[1]
    idx = torch.zeros(B, M, 1, dtype=torch.int32)
    idx_mask = torch.zeros(B, M, 1, dtype=torch.int32)
    lib().cl3d_oracle_nearest_query(_fp(query_xyz), _fp(support_xyz), _ip(query_mask), _ip(support_mask),
                                    B, N, M, _ip(idx), _ip(idx_mask))
    return [idx, idx_mask]


def group_points(points, idx):
    """group_points.cpp:17-40: (B,C,N) f32, (B,M,K) i32 -> (B,C,M,K) f32"""
    _chk(points, "points", torch.float32)
    _chk(idx, "idx", torch.int32)
    B, C, N = points.shape
    _, M, K = idx.shape
    out = torch.zeros(B, C, M, K, dtype=torch.float32)
    lib().cl3d_oracle_group_points(_fp(points), _ip(idx), B, C, N, M, K, _fp(out))
    return out


def group_points_grad(grad_out, idx, n):
    """group_points.cpp:42-65: (B,C,M,K) f32, (B,M,K) i32, n -> (B,C,n) f32"""
    _chk(grad_out, "grad_out", torch.float32)
    _chk(idx, "idx", torch.int32)
    B, C, M, K = grad_out.shape
    out = torch.zeros(B, C, n, dtype=torch.float32)
    lib().cl3d_oracle_group_points_grad(_fp(grad_out), _ip(idx), B, C, int(n), M, K, _fp(out))
    return out


def masked_grid_subsampling(points, mask, nsamples, sampleDl):
    """masked_grid_subsampling.cpp -> [sub_xyz (B,m,3) f32, sub_mask (B,m) i32]"""
    _chk(points, "points", torch.float32)
    _chk(mask, "mask", torch.int32)
    B, n, _ = points.shape
    out = torch.zeros(B, nsamples, 3, dtype=torch.float32)
    out_mask = torch.zeros(B, nsamples, dtype=torch.int32)
    lib().cl3d_oracle_grid_subsample(_fp(points), _ip(mask), B, n, int(nsamples), ctypes.c_float(sampleDl),
                                     _fp(out), _ip(out_mask))
    return [out, out_mask]
