"""Kernel-point dispositions for the PseudoGrid operator (init-time only, not on the hot path).

The reference obtains its `K_points` buffer from a KPConv-style potential optimisation with an on-disk
cache and a rank-0 spin-wait (/root/reference/pytorch/models/utlis.py:10-284).  That machinery is out of
scope; checkpoints carry `K_points` in the state dict and the parity tests copy it from
the reference module.  For a freshly constructed module this file provides an independent, deterministic
generator with the same contract: `num_kpoints` points inside a ball of the given radius, mutually
repelling, with the first point at the centre when fixed == 'center' (or the first three on the vertical
axis when fixed == 'verticals').
"""
import numpy as np


def create_kernel_points(radius, num_kpoints, num_kernels=1, dimension=3, fixed="center", seed=0, iters=3000):
    """-> float32 array (num_kernels, num_kpoints, dimension)"""
    if dimension != 3:
        raise ValueError("Unsupported dimpension of kernel : " + str(dimension))
    rng = np.random.RandomState(seed)
    out = np.zeros((num_kernels, num_kpoints, 3), dtype=np.float64)
    for kern in range(num_kernels):
        # start: points uniformly inside the unit ball
        p = rng.normal(size=(num_kpoints, 3))
        p /= np.linalg.norm(p, axis=1, keepdims=True) + 1e-12
        p *= rng.rand(num_kpoints, 1) ** (1.0 / 3.0)
        step = 0.02
        for _ in range(iters):
            if fixed == "center":
                p[0] = 0.0
            elif fixed == "verticals":
                p[:3, :2] = 0.0
                p[0] = 0.0
            d = p[:, None, :] - p[None, :, :]                       # (n,n,3)
            r2 = np.sum(d * d, axis=-1) + np.eye(num_kpoints)       # avoid /0 on the diagonal
            rep = np.sum(d / (r2[..., None] ** 1.5 + 1e-9), axis=1)  # Coulomb repulsion between points
            att = -2.0 * p                                           # harmonic attraction to the centre
            g = rep * (1.0 / num_kpoints) + att * 0.5
            n = np.linalg.norm(g, axis=1, keepdims=True)
            g = np.where(n > 1.0, g / n, g)                          # clip
            p = p + step * g
            step *= 0.999
        if fixed == "center":
            p[0] = 0.0
        # rescale so the farthest point sits at ~ the requested radius * 2/3 (points live inside the ball)
        far = np.max(np.linalg.norm(p, axis=1))
        p = p / (far + 1e-12) * (2.0 / 3.0)
        out[kern] = p * radius
    return out.astype(np.float32)


def weight_variable(size, rng=None):
    """Truncated-normal init of PseudoGrid.kernel_weights: std = sqrt(2/size[-1]), values beyond 2 std are
    zeroed (contract of the reference's models/utlis.py:297-303)."""
    import torch
    rng = rng or np.random
    std = np.sqrt(2.0 / size[-1])
    w = rng.normal(scale=std, size=size)
    w[np.abs(w) > 2 * std] = 0
    return torch.nn.Parameter(torch.from_numpy(w).float(), requires_grad=True)
