"""CPU model of the cell-ring nearest query (closerlook3d_b200/csrc/neighbors.cu: nearest_query_grid_kernel) in numpy
float32, checked against the reference's index-order scan (masked_nearest_query_gpu.cu:35-52: start value 100, strict
`<`, i.e. the first minimum in index order) on many random clouds.

What this pins is the ALGORITHM -- the grid sizing without a radius, the shell enumeration of a Chebyshev ring, and
above all the stopping bound with its rounding margin -- over far more configurations (clustered, planar, duplicated
points, queries outside the bounding box, coordinates far from the origin, tiny extents) than the GPU parity tests
visit.  The arithmetic mirrors the kernel statement by statement; keep the two in step.
"""
import os
import zlib

import numpy as np
import pytest

F = np.float32


def _d2(q, p):
    """ref_d2 of common.cuh: t = dy*dy; t = fma(dx,dx,t); t = fma(dz,dz,t), d = query - support"""
    dx, dy, dz = F(q[0] - p[0]), F(q[1] - p[1]), F(q[2] - p[2])
    t = F(dy * dy)
    t = F(np.float64(dx) * np.float64(dx) + np.float64(t))
    t = F(np.float64(dz) * np.float64(dz) + np.float64(t))
    return t


def _scan(q, pts):
    best_d, best_i = F(100.0), -1
    for i, p in enumerate(pts):
        d = _d2(q, p)
        if d < best_d:
            best_d, best_i = d, i
    return best_i


def _grid(pts, cell_cap):
    """make_grid_params with radius = 0 + the counting sort of grid_build_fused_kernel"""
    mn, mx = pts.min(0), pts.max(0)
    ex, ey, ez = (F(mx[a] - mn[a]) for a in range(3))
    h = F(1e-20)
    g = (1, 1, 1)
    for _ in range(400):
        f = [min(F(e / h), F(510.0)) for e in (ex, ey, ez)]
        t = [int(v) + 1 for v in f]
        if t[0] * t[1] * t[2] <= cell_cap and ex / h < 511 and ey / h < 511 and ez / h < 511:
            g = tuple(t)
            break
        h = F(h * F(1.25))
    inv_h = F(F(1.0) / h)
    o = mn.astype(F)

    def coord(x, a):
        c = int(F(F(x - o[a]) * inv_h))          # C cast: truncation toward zero
        return 0 if c < 0 else (g[a] - 1 if c >= g[a] else c)

    cells = [[] for _ in range(g[0] * g[1] * g[2])]
    for i, p in enumerate(pts):
        cells[coord(p[0], 0) + g[0] * (coord(p[1], 1) + g[1] * coord(p[2], 2))].append(i)
    return o, inv_h, g, cells, coord


def _rings(q, pts, grid):
    o, inv_h, g, cells, coord = grid
    h = F(F(1.0) / inv_h)
    c = [coord(q[a], a) for a in range(3)]
    best_d, best_i = F(100.0), -1
    visited = 0

    def scan_cells(base, x0, x1):
        nonlocal best_d, best_i, visited
        for x in range(x0, x1 + 1):
            for i in cells[base + x]:
                visited += 1
                d = _d2(q, pts[i])
                if d < best_d or (d == best_d and i < best_i):
                    best_d, best_i = d, i

    rmax = max(max(c[a], g[a] - 1 - c[a]) for a in range(3))
    for r in range(rmax + 1):
        lo = [max(c[a] - r, 0) for a in range(3)]
        hi = [min(c[a] + r, g[a] - 1) for a in range(3)]
        for z in range(lo[2], hi[2] + 1):
            for y in range(lo[1], hi[1] + 1):
                base = g[0] * (y + g[1] * z)
                if abs(z - c[2]) == r or abs(y - c[1]) == r:
                    scan_cells(base, lo[0], hi[0])
                else:
                    if c[0] - r >= 0:
                        scan_cells(base, c[0] - r, c[0] - r)
                    if c[0] + r <= g[0] - 1:
                        scan_cells(base, c[0] + r, c[0] + r)
        bound = F(3.0e38)
        for a in range(3):
            if c[a] - r > 0:
                face = F(o[a] + F(F(c[a] - r) * h))
                bound = min(bound, F(F(q[a] - face) - F(F(0.01) * h + F(2e-6) * F(abs(q[a]) + abs(face)))))
            if c[a] + r < g[a] - 1:
                face = F(o[a] + F(F(c[a] + r + 1) * h))
                bound = min(bound, F(F(face - q[a]) - F(F(0.01) * h + F(2e-6) * F(abs(q[a]) + abs(face)))))
        if bound > 0:
            if best_i >= 0 and F(np.sqrt(best_d)) < bound:
                break
            if bound > F(10.01):
                break
    return best_i, visited


def _cloud(rng, kind, n, offset):
    if kind == "uniform":
        p = rng.random((n, 3))
    elif kind == "shell":                      # surface-like: most cells empty
        d = rng.standard_normal((n, 3))
        p = d / np.linalg.norm(d, axis=1, keepdims=True) * (0.5 + 0.01 * rng.standard_normal((n, 1)))
    elif kind == "plane":                      # zero extent along one axis
        p = np.concatenate([rng.random((n, 2)), np.full((n, 1), 0.25)], 1)
    elif kind == "clusters":
        centres = rng.random((4, 3)) * 3
        p = centres[rng.integers(0, 4, n)] + 0.02 * rng.standard_normal((n, 3))
    elif kind == "tiny":                       # extent ~1e-4 around a far offset
        p = 1e-4 * rng.random((n, 3))
    else:
        raise ValueError(kind)
    p = (p + offset).astype(F)
    m = n // 10
    if m:
        p[:m] = p[n - m:]                      # duplicated points: distance ties -> the smaller index wins
    return p


@pytest.mark.parametrize("kind", ["uniform", "shell", "plane", "clusters", "tiny"])
@pytest.mark.parametrize("offset", [0.0, 25.0, -300.0])
def test_ring_walk_equals_index_order_scan(kind, offset):
    rng = np.random.default_rng(zlib.crc32(f"{kind}{offset}".encode()) + int(os.environ.get("CL3D_FUZZ_SEED", "0")))
    total_visited = total_pairs = 0
    for trial in range(int(os.environ.get("CL3D_FUZZ_TRIALS", "6"))):
        n = int(rng.integers(3, 260))
        pts = _cloud(rng, kind, n, offset)
        grid = _grid(pts, max(4 * n, 4096))
        span = float(np.abs(pts - pts.mean(0)).max()) + 1e-6
        qs = [pts[int(rng.integers(0, n))] for _ in range(6)]                              # on support points
        qs += [(pts.mean(0) + span * 1.5 * (rng.random(3) - 0.5)).astype(F) for _ in range(14)]   # around the cloud
        qs += [(pts.mean(0) + span * 6.0 * (rng.random(3) - 0.5)).astype(F) for _ in range(6)]    # far outside the box
        for q in qs:
            want = _scan(q, pts)
            got, visited = _rings(q, pts, grid)
            assert got == want, (kind, offset, trial, n, q, want, got)
            total_visited += visited
            total_pairs += n
    assert total_visited <= total_pairs          # never worse than the scan; far fewer on spread-out clouds


def test_beyond_the_start_value_returns_minus_one():
    # every support farther than sqrt(100) = 10 from the query: the reference keeps min_idx = -1
    pts = (np.random.default_rng(1).random((50, 3)) + 100.0).astype(F)
    q = np.zeros(3, F)
    assert _scan(q, pts) == -1
    assert _rings(q, pts, _grid(pts, 4096))[0] == -1
