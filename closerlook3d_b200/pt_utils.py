"""Drop-in replacement for the reference's `pt_utils` module
(/root/reference/pytorch/ops/pt_custom_ops/pt_utils.py): same public names, constructor / forward
signatures, return shapes and autograd behaviour, running on libcl3d's sm_100a kernels.

    reference name (pt_utils.py line)            here
    grouping_operation            (:64)          GroupingOperation.apply      -> cl3d_group_points[_grad]
    masked_ordered_ball_query     (:80)          MaskedOrderedBallQuery.apply -> cl3d_ball_query (grid hash)
    masked_nearest_query          (:95)          MaskedNearestQuery.apply     -> cl3d_nearest_query
    masked_grid_subsampling       (:111)         MaskedGridSubsampling.apply  -> cl3d_grid_subsample
    MaskedQueryAndGroup           (:114)         same forward contract (materialising compat path)
    MaskedNearestQueryAndGroup    (:147)
    MaskedMaxPool                 (:179)         fused: no (B,C,M,K) tensor
    MaskedUpsample                (:205)         nearest: fused gather of one row per query

The fused LocalAggregation operators (local_aggregation_operators.py in this package) do NOT go through
MaskedQueryAndGroup: they share its neighbour search via `neighbors()` below, which also caches the
neighbour lists so that the duplicate queries of a backbone forward (la1/btnk1, MaskedMaxPool + the strided
block's LocalAggregation: 5 of 14) are searched once.
"""
import collections

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import ops


# --------------------------------------------------------------------------------------------------
# neighbour lists + cache
# --------------------------------------------------------------------------------------------------
class NeighborList:
    """idx (B,M,K) i32, ncount (B,M) i32 [+ idx_mask (B,M,K) i32 on demand, CSR lists on demand].

    The search (and, when a backward will follow, the transposed lists) is enqueued on a side stream so that it
    overlaps the layout change / per-point product of the caller; `wait()` / `csr()` make the current stream
    wait for it.  Every use of the side stream starts with side.wait_stream(current), so memory allocated there
    is never recycled under a pending consumer."""

    def __init__(self, idx, ncount, idx_mask, N):
        self.idx, self.ncount, self._idx_mask, self.N = idx, ncount, idx_mask, N
        self._csr = None
        self._csr_all = None
        self.parts = None          # [(b0, b1, search-done event)] when the batch was searched in parts
        self.event = None
        self._csr_event = None
        self._csr_all_event = None

    def wait(self):
        if self.event is not None:
            torch.cuda.current_stream().wait_event(self.event)
        return self

    @property
    def idx_mask(self):
        if self._idx_mask is None:
            raise RuntimeError("idx_mask was not requested for this neighbour list")
        return self._idx_mask

    def _build(self, all_slots):
        if all_slots:
            full = torch.full_like(self.ncount, self.idx.shape[2])
            return ops.build_csr(self.idx, full, self.N)
        return ops.build_csr(self.idx, self.ncount, self.N)

    def prefetch_csr(self, all_slots=False):
        """enqueue the transposed-list build on the side stream right behind the search"""
        if (self._csr_all if all_slots else self._csr) is not None:
            return
        side = _side_stream(self.idx.device)
        if self.event is not None:
            side.wait_event(self.event)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            csr = self._build(all_slots)
            ev = torch.cuda.Event()
            ev.record(side)
        if all_slots:
            self._csr_all, self._csr_all_event = csr, ev
        else:
            self._csr, self._csr_event = csr, ev

    def csr(self):
        """transposed lists over the COUNTED slots (k < ncount): avg/sum families"""
        if self._csr is None:
            self.wait()
            self._csr = self._build(False)
        elif self._csr_event is not None:
            torch.cuda.current_stream().wait_event(self._csr_event)
        return self._csr

    def csr_all_slots(self):
        """transposed lists over ALL K slots (BatchNorm2d of PointWiseMLP sees every slot)"""
        if self._csr_all is None:
            self.wait()
            self._csr_all = self._build(True)
        elif self._csr_all_event is not None:
            torch.cuda.current_stream().wait_event(self._csr_all_event)
        return self._csr_all


_SIDE = {}
overlap_enabled = True  # run the neighbour search on a side stream (fork/join), see NeighborList
# Search the clouds of a batch in this many parts (clouds are independent).  The fused operators start the forward of
# part 0 as soon as ITS search is done, so the issue-bound search of part 1 runs beside the L1-bound aggregation of
# part 0 instead of in front of it.  1 = off.
batch_parts = 1   # measured at c3 (profiles/RESULTS_r2.md): 2 parts 1.154 ms/step vs 1.107 -- half-batch kernels lose more to tail effects than the overlap wins


def _side_stream(device, priority=0):
    """per-device helper stream; priority=-1: a second, high-priority one (its CTAs are scheduled first when it
    and the main stream both have a kernel pending -- used for short latency-bound passes beside a long one)"""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), priority)
    st = _SIDE.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device, priority=priority)
        _SIDE[key] = st
    return st


_CACHE = collections.OrderedDict()
_CACHE_SIZE = 8
cache_enabled = True
cache_stats = {"hit": 0, "miss": 0}


def _key(t):
    return (t.data_ptr(), tuple(t.shape), t._version, t.device.index)


def clear_neighbor_cache():
    _CACHE.clear()


def neighbors(query_xyz, support_xyz, query_mask, support_mask, radius, nsample, need_mask=False, overlap=None,
              csr=None):
    """Ball-query neighbour list for (query, support, radius, nsample), cached on tensor identity
    (data pointer + version; the cache keeps the key tensors alive so pointers cannot be recycled).
    With overlap (default: pt_utils.overlap_enabled) the search runs on a side stream; consumers call
    nl.wait() (the fused operators do) before touching nl.idx / nl.ncount.
    csr = "counted" | "all": a backward will follow -- the transposed lists are built in the same call as the
    search (ranks taken in its emit phase), not by a separate count / scan / fill afterwards."""
    key = (_key(query_xyz), _key(support_xyz), _key(query_mask), _key(support_mask), float(radius), int(nsample))
    if cache_enabled:
        hit = _CACHE.get(key)
        if hit is not None and (hit[0]._idx_mask is not None or not need_mask):
            _CACHE.move_to_end(key)
            cache_stats["hit"] += 1
            return hit[0]
    cache_stats["miss"] += 1
    overlap = overlap_enabled if overlap is None else overlap
    lists = None
    if overlap:
        cur = torch.cuda.current_stream()
        side = _side_stream(query_xyz.device)
        side.wait_stream(cur)
        B = query_xyz.shape[0]
        nparts = max(1, min(int(batch_parts), B)) if csr else 1   # parts only pay off when a fused forward follows
        parts = None
        with torch.cuda.stream(side):
            if nparts == 1:
                ev = torch.cuda.Event()      # search done: idx / ncount final (the forward kernels wait for this one)
                res = ops.ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample,
                                     want_mask=need_mask, want_ncount=True, csr=csr,
                                     after_search=(lambda: ev.record(side)) if csr else None)
                if not csr:
                    ev.record(side)
                idx, idx_mask, ncount = res[:3]
                lists = res[3] if csr else None
            else:
                M, N, K, dev = query_xyz.shape[1], support_xyz.shape[1], int(nsample), query_xyz.device
                idx = torch.empty(B, M, K, dtype=torch.int32, device=dev)
                idx_mask = torch.empty(B, M, K, dtype=torch.int32, device=dev) if need_mask else None
                ncount = torch.empty(B, M, dtype=torch.int32, device=dev)
                lists = (torch.empty(B, N + 1, dtype=torch.int32, device=dev),
                         torch.empty(B, M * K, dtype=torch.int32, device=dev))
                parts = []
                for h in range(nparts):
                    b0, b1 = (B * h) // nparts, (B * (h + 1)) // nparts
                    evh = torch.cuda.Event()
                    out = dict(idx=idx[b0:b1], idx_mask=idx_mask[b0:b1] if need_mask else None, ncount=ncount[b0:b1],
                               off=lists[0][b0:b1], ent=lists[1][b0:b1])
                    ops.ball_query(query_xyz[b0:b1], support_xyz[b0:b1], query_mask[b0:b1], support_mask[b0:b1], radius,
                                   nsample, want_mask=need_mask, want_ncount=True, csr=csr, out=out,
                                   after_search=(lambda e=evh: e.record(side)))
                    parts.append((b0, b1, evh))
                ev = parts[-1][2]            # every part searched (stream order)
            ev_lists = None
            if csr:                          # lists done: only the backward waits for this one
                ev_lists = torch.cuda.Event()
                ev_lists.record(side)
        nl = NeighborList(idx, ncount, idx_mask, support_xyz.shape[1])
        nl.event = ev
        nl.parts = parts
    else:
        res = ops.ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample,
                             want_mask=need_mask, want_ncount=True, csr=csr)
        idx, idx_mask, ncount = res[:3]
        lists = res[3] if csr else None
        nl = NeighborList(idx, ncount, idx_mask, support_xyz.shape[1])
    if lists is not None:
        lists_event = ev_lists if overlap else None    # in-order on the current stream without overlap
        if csr == "all":
            nl._csr_all, nl._csr_all_event = lists, lists_event
        else:
            nl._csr, nl._csr_event = lists, lists_event
    if cache_enabled:
        _CACHE[key] = (nl, (query_xyz, support_xyz, query_mask, support_mask))
        while len(_CACHE) > _CACHE_SIZE:
            _CACHE.popitem(last=False)
    return nl


# --------------------------------------------------------------------------------------------------
# autograd functions with the reference's names
# --------------------------------------------------------------------------------------------------
class GroupingOperation(Function):
    """pt_utils.py:16-61"""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.for_backwards = (idx, features.size(2))
        return ops.group_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        return ops.group_points_grad(grad_out.contiguous(), idx, N), None


grouping_operation = GroupingOperation.apply


class MaskedOrderedBallQuery(Function):
    """pt_utils.py:67-77"""

    @staticmethod
    def forward(ctx, radius, nsample, query_xyz, support_xyz, query_mask, support_mask):
        nl = neighbors(query_xyz, support_xyz, query_mask, support_mask, radius, nsample, need_mask=True).wait()
        ctx.mark_non_differentiable(nl.idx, nl.idx_mask)
        return nl.idx, nl.idx_mask

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None, None, None


masked_ordered_ball_query = MaskedOrderedBallQuery.apply


class MaskedNearestQuery(Function):
    """pt_utils.py:83-92; returns (B,M,1) tensors like the reference"""

    @staticmethod
    def forward(ctx, query_xyz, support_xyz, query_mask, support_mask):
        idx, idx_mask = ops.nearest_query(query_xyz, support_xyz, query_mask, support_mask)
        idx, idx_mask = idx.unsqueeze(-1), idx_mask.unsqueeze(-1)
        ctx.mark_non_differentiable(idx, idx_mask)
        return idx, idx_mask

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None


masked_nearest_query = MaskedNearestQuery.apply


class MaskedGridSubsampling(Function):
    """pt_utils.py:98-108"""

    @staticmethod
    def forward(ctx, xyz, mask, npoint, sampleDl):
        sub_xyz, sub_mask = ops.grid_subsample(xyz, mask, npoint, sampleDl)
        ctx.mark_non_differentiable(sub_xyz, sub_mask)
        return sub_xyz, sub_mask

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None


masked_grid_subsampling = MaskedGridSubsampling.apply


# --------------------------------------------------------------------------------------------------
# modules
# --------------------------------------------------------------------------------------------------
class MaskedQueryAndGroup(nn.Module):
    """pt_utils.py:114-144 (materialising compatibility path: returns the (B,C,M,K) tensors)."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz

    def forward(self, query_xyz, support_xyz, query_mask, support_mask, features=None):
        idx, idx_mask = masked_ordered_ball_query(self.radius, self.nsample, query_xyz, support_xyz, query_mask,
                                                  support_mask)
        xyz_trans = support_xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)  # (B,3,M,K)
        grouped_xyz = grouped_xyz - query_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz = grouped_xyz / self.radius
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        if self.ret_grouped_xyz:
            return new_features, grouped_xyz, idx_mask
        return new_features, idx_mask


class MaskedNearestQueryAndGroup(nn.Module):
    """pt_utils.py:147-176"""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False):
        super().__init__()
        self.use_xyz, self.ret_grouped_xyz, self.normalize_xyz = use_xyz, ret_grouped_xyz, normalize_xyz

    def forward(self, query_xyz, support_xyz, query_mask, support_mask, features=None):
        idx, idx_mask = masked_nearest_query(query_xyz, support_xyz, query_mask, support_mask)
        xyz_trans = support_xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)
        grouped_xyz = grouped_xyz - query_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:  # the reference reads self.radius here, which this class never defines (:160)
            raise AttributeError("'MaskedNearestQueryAndGroup' object has no attribute 'radius'")
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        if self.ret_grouped_xyz:
            return new_features, grouped_xyz, idx_mask
        return new_features, idx_mask


class _GatherMax(Function):
    """out[b,c,q] = max_k f[b,c,idx[b,q,k]]  (fused MaskedMaxPool body: pt_utils.py:195-201)"""

    @staticmethod
    def forward(ctx, features, idx):
        out, arg = ops.gather_max(features, idx)
        ctx.save_for_backward(idx, arg)
        ctx.N = features.shape[2]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, arg = ctx.saved_tensors
        return ops.gather_max_grad(grad_out.contiguous(), idx, arg, ctx.N), None


class MaskedMaxPool(nn.Module):
    """pt_utils.py:179-202"""

    def __init__(self, npoint, radius, nsample, sampleDl):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.sampleDl = npoint, radius, nsample, sampleDl

    def forward(self, xyz, mask, features):
        sub_xyz, sub_mask = masked_grid_subsampling(xyz, mask, self.npoint, self.sampleDl)
        sub_xyz = sub_xyz.contiguous()
        sub_mask = sub_mask.contiguous()
        nl = neighbors(sub_xyz, xyz, sub_mask, mask, self.radius, self.nsample).wait()
        sub_features = _GatherMax.apply(features.contiguous(), nl.idx)
        return sub_xyz, sub_mask, sub_features


class MaskedUpsample(nn.Module):
    """pt_utils.py:205-227"""

    def __init__(self, radius, nsample, mode='nearest'):
        super().__init__()
        self.radius, self.nsample, self.mode = radius, nsample, mode

    def forward(self, up_xyz, xyz, up_mask, mask, features):
        if self.mode == 'nearest':
            idx, _ = masked_nearest_query(up_xyz, xyz, up_mask, mask)  # (B,M,1)
            return grouping_operation(features, idx)[..., 0].contiguous()
        elif self.mode == 'max':
            nl = neighbors(up_xyz, xyz, up_mask, mask, self.radius, self.nsample).wait()
            return _GatherMax.apply(features.contiguous(), nl.idx)
        raise NotImplementedError(f"mode:{self.mode} not supported in MaskedUpsample")
