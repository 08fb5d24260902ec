"""debug: PointWiseMLP smoke case, variations"""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from closerlook3d_b200 import ops, synth, pt_utils
from closerlook3d_b200.config import la_config
from closerlook3d_b200.local_aggregation_operators import LocalAggregation
import oracle; oracle.build()
from oracle import ext as oext, la_oracle
dev = torch.device("cuda:0")
la_oracle.GPU_SCALAR_DIVISION = True

def run(seed_m, seed_d, B, N, K, C, overlap=True, pre_query=False, tag=""):
    pt_utils.overlap_enabled = overlap
    la, over = "pointwisemlp", dict(pointwisemlp=dict(feature_type="dp_fi_df", num_mlps=1, reduction="max"))
    cfg = la_config(la, **over)
    torch.manual_seed(seed_m); np.random.seed(seed_m)
    r = synth.ball_radius(N, K)
    mod = LocalAggregation(C, C, r, K, cfg)
    sd = copy.deepcopy(mod.state_dict())
    d = synth.make_cloud_batch(B, N, C, seed_d)
    xyz, mask, feats = d["xyz"], d["mask"], d["features"]
    gout = torch.randn(B, C, N, generator=torch.Generator().manual_seed(3))
    if pre_query:
        ops.ball_query(xyz.to(dev), xyz.to(dev), mask.to(dev), mask.to(dev), r, K)
    orc = la_oracle.OracleLocalAggregation(oext, la, C, C, r, K, cfg, sd)
    f_ref = feats.clone().requires_grad_(True)
    o_ref = orc(xyz, xyz, mask, mask, f_ref)
    mod = mod.to(dev).train()
    f = feats.to(dev).requires_grad_(True)
    out = mod(xyz.to(dev), xyz.to(dev), mask.to(dev), mask.to(dev), f)
    keep = ~((out.detach().cpu() > 0) != (o_ref.detach() > 0))
    (o_ref * gout * keep).sum().backward()
    (out * (gout * keep).to(dev)).sum().backward()
    torch.cuda.synchronize()
    e_out = float((out.detach().cpu() - o_ref.detach()).abs().max()) / max(1.0, float(o_ref.detach().abs().max()))
    dg = (f.grad.cpu() - f_ref.grad).abs()
    e_g = float(dg.max()) / max(1.0, float(f_ref.grad.abs().max()))
    nbad = int((dg > 1e-4 * float(f_ref.grad.abs().max())).sum())
    loc = np.unravel_index(int(dg.argmax()), dg.shape)
    pg = {k: v for k, v in orc.grads().items()}
    epar = {n: float((p.grad.cpu() - pg[n[len("local_aggregation_operator."):]]).abs().max()) /
            max(1.0, float(pg[n[len("local_aggregation_operator."):]].abs().max())) for n, p in mod.named_parameters()}
    if nbad:
        bad = dg > 1e-4 * float(f_ref.grad.abs().max())
        print("   bad per cloud", bad.sum((1, 2)).tolist(), "bad channels", sorted(set(bad.nonzero()[:, 1].tolist()))[:20],
              "bad points", sorted(set(bad.nonzero()[:, 2].tolist()))[:20], "n bad points", len(set(bad.nonzero()[:, 2].tolist())))
        print("   param errs", {k[-20:]: f"{v:.1e}" for k, v in epar.items()})
        idx, _, nc = ops.ball_query(xyz.to(dev), xyz.to(dev), mask.to(dev), mask.to(dev), r, K)
        print("   ncount min/max", int(nc.min()), int(nc.max()))
        # which queries reference the bad points?
        bp = sorted(set(bad.nonzero()[:, 2].tolist()))[:3]
        for j in bp:
            qs = (idx[1] == j).any(-1).nonzero().flatten().tolist()
            print("   point", j, "referenced by", len(qs), "queries; mask", int(mask[1, j]), "first refs", qs[:8])
        gW = [p.grad for n, p in mod.named_parameters() if n.endswith("conv0.0.weight")][0].cpu()
        rW = pg["mlps.conv0.0.weight"]
        dW = (gW - rW).abs().reshape(gW.shape[0], -1)
        print("   dW err by out-channel (top)", torch.topk(dW.max(1)[0], 5), "by in-col", torch.topk(dW.max(0)[0], 5))
    print(f"{tag:28s} seeds({seed_m},{seed_d}) B{B} N{N} K{K} C{C} ovl={overlap}: out {e_out:.1e} grad {e_g:.1e} nbad {nbad} at {loc} flips {int((~keep).sum())} params {max(epar.values()):.1e}")

run(11, 1005, 2, 1024, 16, 72, tag="smoke case")
run(11, 1005, 2, 1024, 16, 72, pre_query=True, tag="smoke case + pre query")
run(11, 1005, 2, 1024, 16, 72, overlap=False, tag="no overlap")
for s in range(5):
    run(20 + s, 2000 + s, 2, 1024, 16, 72, tag="other seeds")
run(3090, 3090, 2, 1024, 16, 66, tag="test-suite case")
run(11, 1005, 1, 1024, 16, 72, tag="B=1")
run(11, 1005, 2, 1024, 16, 24, tag="C=24")
