#!/usr/bin/env python
"""cl3d_nearest_query: cell-ring walk vs tile scan at the S3DIS upsampling shapes (CUDA events, 20 runs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from closerlook3d_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
for B, M, N in ((8, 15000, 3750), (8, 40000, 10000), (16, 10000, 2500)):
    g = torch.Generator().manual_seed(1)
    d = torch.randn(B, M, 3, generator=g)
    q = (d / d.norm(dim=-1, keepdim=True) * (0.5 + 0.02 * torch.randn(B, M, 1, generator=g))).to(dev)
    s = q[:, torch.randperm(M, generator=g)[:N]].contiguous()
    qm = torch.ones(B, M, dtype=torch.int32, device=dev)
    sm = torch.ones(B, N, dtype=torch.int32, device=dev)
    res = {}
    for mode in ("grid", "scan"):
        if mode == "scan":
            os.environ["CL3D_NN_BRUTE"] = "1"
        else:
            os.environ.pop("CL3D_NN_BRUTE", None)
        for _ in range(3):
            idx, _m = ops.nearest_query(q, s, qm, sm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            idx, _m = ops.nearest_query(q, s, qm, sm)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 20, idx)
    os.environ.pop("CL3D_NN_BRUTE", None)
    print(f"B={B} M={M} N={N}: cell rings {res['grid'][0]:.4f} ms, tile scan {res['scan'][0]:.4f} ms, "
          f"identical={bool(torch.equal(res['grid'][1], res['scan'][1]))}")
