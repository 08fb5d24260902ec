#!/usr/bin/env python
"""Generate tests/golden/*.pt from the UNMODIFIED reference python modules (container only).

TEST INFRASTRUCTURE ONLY.  Run where /root/reference exists:   python -m oracle.make_golden

For every operator family a small seeded problem is pushed through the reference's own
`models/local_aggregation_operators.LocalAggregation` (imported unmodified by oracle/ref_loader.py, running on
the C restatement of its CUDA ops) and inputs, parameters, neighbour indices, outputs and gradients are
stored.  The fixtures travel to the GPU box, where /root/reference does not exist: tests compare the oracle
(bit-exact) and the CUDA path (1e-5) against them.
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from closerlook3d_b200 import synth  # noqa: E402
from oracle import ext, ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (la_type, overrides, B, N, K, C, M or None)
    "pospool_xyz_avg": ("pospool", dict(pospool=dict(position_embedding="xyz", reduction="avg")), 2, 256, 12, 18, None),
    "pospool_sincos_avg": ("pospool", dict(pospool=dict(position_embedding="sin_cos", reduction="avg")), 2, 256, 12, 24, None),
    "adaptive_weight_dp": ("adaptive_weight", dict(adaptive_weight=dict(weight_type="dp", num_mlps=1, shared_channels=1,
                                                                       reduction="avg")), 2, 256, 12, 24, None),
    "pointwisemlp_dp_fi_df": ("pointwisemlp", dict(pointwisemlp=dict(feature_type="dp_fi_df", num_mlps=1,
                                                                    reduction="max")), 2, 256, 12, 24, None),
    "pseudo_grid": ("pseudo_grid", dict(), 2, 256, 12, 24, None),
    "pospool_xyz_strided": ("pospool", dict(pospool=dict(position_embedding="xyz", reduction="avg")), 2, 320, 10, 18, 96),
}


def main():
    ns = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    for name, (la_type, over, B, N, K, C, M) in CASES.items():
        seed = 4242 + len(name)
        torch.manual_seed(seed)
        import numpy as np
        np.random.seed(seed)
        cfg = ref_loader.make_config(la_type, **over)
        r = synth.ball_radius(N, K)
        ref = ns.lao.LocalAggregation(C, C, r, K, cfg)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for n_, p in ref.named_parameters():
                if "out_transform" in n_ or ".1." in n_:
                    p.copy_((1.0 if n_.endswith("weight") else 0.0) + 0.5 * torch.randn(p.shape, generator=g))
        sd = copy.deepcopy(ref.state_dict())
        d = synth.make_cloud_batch(B, N, C, seed)
        xyz, mask, feats = d["xyz"], d["mask"], d["features"]
        if M is None:
            q, qm = xyz, mask
        else:
            q = (xyz[:, :M] + 0.1 * r * torch.randn(B, M, 3, generator=g)).contiguous()
            qm = torch.ones(B, M, dtype=torch.int32)
            qm[:, M - M // 8:] = 0
        gout = torch.randn(B, C, q.shape[1], generator=g)
        f = feats.clone().requires_grad_(True)
        out = ref(q, xyz, qm, mask, f)
        (out * gout).sum().backward()
        idx, idx_mask = ext.masked_ordered_ball_query(q, xyz, qm, mask, r, K)
        blob = dict(la_type=la_type, overrides=over, B=B, N=N, K=K, C=C, radius=r,
                    query_xyz=q, support_xyz=xyz, query_mask=qm, support_mask=mask, features=feats, grad_out=gout,
                    state_dict=sd, out=out.detach(), grad_features=f.grad,
                    grad_params={k: v.grad for k, v in ref.named_parameters()},
                    state_dict_after={k: v.clone() for k, v in ref.state_dict().items()},
                    idx=idx, idx_mask=idx_mask,
                    made_by="oracle/make_golden.py: unmodified reference LocalAggregation on oracle.ext (CPU)")
        torch.save(blob, os.path.join(OUT, f"la_{name}.pt"))
        print(name, "out", tuple(out.shape), "bytes", os.path.getsize(os.path.join(OUT, f"la_{name}.pt")))


if __name__ == "__main__":
    main()
