"""GPU parity of the "next" rows of SURVEY.md 8f: masked_grid_subsampling (bit-exact), MaskedMaxPool,
MaskedUpsample, the `_ext` drop-in surface and the neighbour-list cache, against the oracle."""
import pytest
import torch

from closerlook3d_b200 import synth

pytestmark = pytest.mark.gpu


def _cloud(B, N, C, seed):
    d = synth.make_cloud_batch(B, N, C, seed)
    return d["xyz"], d["mask"], d["features"]


@pytest.mark.parametrize("B,n,m,dl", [(3, 1024, 300, 0.12), (2, 3000, 1000, 0.07), (2, 500, 600, 0.2),
                                      (4, 15000, 4096, 0.045), (1, 64, 10, 0.5), (2, 2000, 2000, 0.001)])
def test_grid_subsample_bit_exact(cuda, oracle_ext, B, n, m, dl):
    from closerlook3d_b200 import ops
    xyz, mask, _ = _cloud(B, n, 3, 300 + n)
    rs, rm = oracle_ext.masked_grid_subsampling(xyz, mask, m, dl)
    s, sm = ops.grid_subsample(xyz.to(cuda), mask.to(cuda), m, dl)
    assert torch.equal(sm.cpu(), rm)
    assert torch.equal(s.cpu(), rs), f"{(s.cpu() != rs).any(-1).sum().item()} rows differ"


def test_grid_subsample_edge_origin_rounding(cuda, oracle_ext):
    # coordinates straddling multiples of dl, negative coordinates, duplicated points
    from closerlook3d_b200 import ops
    g = torch.Generator().manual_seed(4)
    xyz = (torch.randint(-40, 40, (2, 1500, 3), generator=g).float() * 0.025 + 1e-7 * torch.randn(2, 1500, 3, generator=g))
    xyz = xyz.contiguous()
    mask = torch.ones(2, 1500, dtype=torch.int32)
    mask[1, 1200:] = 0
    for dl in (0.05, 0.1, 0.025):
        rs, rm = oracle_ext.masked_grid_subsampling(xyz, mask, 700, dl)
        s, sm = ops.grid_subsample(xyz.to(cuda), mask.to(cuda), 700, dl)
        assert torch.equal(sm.cpu(), rm) and torch.equal(s.cpu(), rs)


def test_masked_max_pool_matches_oracle(cuda, oracle_ext):
    from closerlook3d_b200.pt_utils import MaskedMaxPool
    from oracle import la_oracle
    B, N, C, npoint, K = 2, 3000, 40, 800, 20
    xyz, mask, feats = _cloud(B, N, C, 12)
    radius, dl = 0.12, 0.08
    f_ref = feats.clone().requires_grad_(True)
    sx, sm, sf = la_oracle.masked_max_pool(oracle_ext, xyz, mask, f_ref, npoint, radius, K, dl)
    gout = torch.randn(sf.shape, generator=torch.Generator().manual_seed(1))
    (sf * gout).sum().backward()
    pool = MaskedMaxPool(npoint, radius, K, dl)
    f = feats.to(cuda).requires_grad_(True)
    x2, m2, f2 = pool(xyz.to(cuda), mask.to(cuda), f)
    (f2 * gout.to(cuda)).sum().backward()
    assert torch.equal(x2.cpu(), sx) and torch.equal(m2.cpu(), sm)
    assert torch.equal(f2.detach().cpu(), sf.detach())          # a max of gathered values: exact
    assert torch.allclose(f.grad.cpu(), f_ref.grad, atol=1e-5)


def test_masked_upsample_nearest_matches_oracle(cuda, oracle_ext):
    from closerlook3d_b200.pt_utils import MaskedUpsample
    from oracle import la_oracle
    B, N, C, M = 2, 700, 24, 2500
    xyz, mask, feats = _cloud(B, N, C, 21)
    g = torch.Generator().manual_seed(2)
    up_xyz = torch.rand(B, M, 3, generator=g)
    up_mask = torch.ones(B, M, dtype=torch.int32)
    f_ref = feats.clone().requires_grad_(True)
    ref = la_oracle.masked_upsample_nearest(oracle_ext, up_xyz, xyz, up_mask, mask, f_ref)
    gout = torch.randn(ref.shape, generator=g)
    (ref * gout).sum().backward()
    up = MaskedUpsample(0.1, 16, mode="nearest")
    f = feats.to(cuda).requires_grad_(True)
    out = up(up_xyz.to(cuda), xyz.to(cuda), up_mask.to(cuda), mask.to(cuda), f)
    (out * gout.to(cuda)).sum().backward()
    assert torch.equal(out.detach().cpu(), ref.detach())
    assert torch.allclose(f.grad.cpu(), f_ref.grad, atol=1e-5)


def test_ext_surface_and_compat_grouper(cuda, oracle_ext):
    """the five `_ext` names on libcl3d + the materialising MaskedQueryAndGroup contract (pt_utils.py:121-144)"""
    from closerlook3d_b200 import ext
    from closerlook3d_b200.pt_utils import MaskedQueryAndGroup
    from oracle import la_oracle
    B, N, C, K = 2, 900, 12, 10
    xyz, mask, feats = _cloud(B, N, C, 33)
    r = synth.ball_radius(N, K)
    gx, gm, gf = xyz.to(cuda), mask.to(cuda), feats.to(cuda)
    idx, idx_mask = ext.masked_ordered_ball_query(gx, gx, gm, gm, r, K)
    ridx, rmask = oracle_ext.masked_ordered_ball_query(xyz, xyz, mask, mask, r, K)
    assert torch.equal(idx.cpu(), ridx) and torch.equal(idx_mask.cpu(), rmask)
    assert torch.equal(ext.group_points(gf, idx).cpu(), oracle_ext.group_points(feats, ridx))
    n1, m1 = ext.masked_nearest_query(gx, gx, gm, gm)
    assert n1.shape == (B, N, 1) and m1.shape == (B, N, 1)
    s1, sm1 = ext.masked_grid_subsampling(gx, gm, 200, 0.15)
    assert s1.shape == (B, 200, 3) and sm1.dtype == torch.int32
    grouper = MaskedQueryAndGroup(r, K, use_xyz=False, ret_grouped_xyz=True, normalize_xyz=True)
    nf, gxyz, gmask = grouper(gx, gx, gm, gm, gf)
    rf, rxyz, rm2, _ = la_oracle.query_and_group(oracle_ext, xyz, xyz, mask, mask, feats, r, K, True)
    assert torch.equal(nf.cpu(), rf) and torch.equal(gmask.cpu(), rm2)
    assert torch.allclose(gxyz.cpu(), rxyz, atol=1e-6)


def test_neighbor_cache_reuses_duplicate_queries(cuda):
    from closerlook3d_b200 import pt_utils
    xyz, mask, _ = _cloud(2, 2500, 3, 44)
    gx, gm = xyz.to(cuda), mask.to(cuda)
    pt_utils.clear_neighbor_cache()
    pt_utils.cache_enabled = True
    h0, m0 = pt_utils.cache_stats["hit"], pt_utils.cache_stats["miss"]
    a = pt_utils.neighbors(gx, gx, gm, gm, 0.1, 16)
    b = pt_utils.neighbors(gx, gx, gm, gm, 0.1, 16)          # la1 / btnk1 style duplicate
    c = pt_utils.neighbors(gx, gx, gm, gm, 0.2, 16)          # different radius -> new search
    assert a is b and c is not a
    assert pt_utils.cache_stats["hit"] == h0 + 1 and pt_utils.cache_stats["miss"] == m0 + 2
    gx.mul_(1.0)                                             # in-place edit bumps the version -> no stale hit
    d = pt_utils.neighbors(gx, gx, gm, gm, 0.1, 16)
    assert d is not a
    pt_utils.clear_neighbor_cache()
