"""debug helper (GPU): fused vs oracle, per-tensor errors, no asserts"""
import copy, sys, torch, numpy as np
sys.path.insert(0, '.')
from closerlook3d_b200 import synth
from closerlook3d_b200.config import la_config
from closerlook3d_b200.local_aggregation_operators import LocalAggregation
import oracle; oracle.build()
from oracle import ext as oext, la_oracle

def rel(a, b):
    return float((a - b).abs().max()) / max(1.0, float(b.abs().max()))

def case(la_type, over, B, N, K, C, seed, M=None, radius=None, noise=0.1):
    cfg = la_config(la_type, **over)
    torch.manual_seed(seed); np.random.seed(seed)
    r = synth.ball_radius(N, K) if radius is None else radius
    mod = LocalAggregation(C, C, r, K, cfg)
    sd = copy.deepcopy(mod.state_dict())
    d = synth.make_cloud_batch(B, N, C, seed)
    xyz, mask, feats = d['xyz'], d['mask'], d['features']
    if M is None: q, qm = xyz, mask
    else:
        g = torch.Generator().manual_seed(seed + 1)
        q = (xyz[:, :M] + noise * r * torch.randn(B, M, 3, generator=g)).contiguous()
        qm = torch.ones(B, M, dtype=torch.int32); qm[:, M - M // 8:] = 0
    gout = torch.randn(B, C, q.shape[1], generator=torch.Generator().manual_seed(seed + 2))
    orc = la_oracle.OracleLocalAggregation(oext, la_type, C, C, r, K, cfg, sd)
    f_ref = feats.clone().requires_grad_(True)
    o_ref = orc(q, xyz, qm, mask, f_ref); (o_ref * gout).sum().backward()
    idx_ref, m_ref = oext.masked_ordered_ball_query(q, xyz, qm, mask, r, K)
    print(f"--- {la_type} {over} B{B} N{N} K{K} C{C} M{M}: oracle nan={bool(torch.isnan(o_ref).any())} min cnt={int(m_ref.sum(-1)[qm>0].min())}")
    dev = torch.device('cuda:0')
    mod = mod.to(dev)
    f = feats.to(dev).requires_grad_(True)
    out = mod(q.to(dev), xyz.to(dev), qm.to(dev), mask.to(dev), f); (out * gout.to(dev)).sum().backward()
    torch.cuda.synchronize()
    print("   out", rel(out.detach().cpu(), o_ref.detach()), "nan", bool(torch.isnan(out).any()), " gradf", rel(f.grad.cpu(), f_ref.grad))
    # where is gradf wrong
    dg = (f.grad.cpu() - f_ref.grad).abs()
    bad = (dg > 1e-4 * f_ref.grad.abs().max()).nonzero()
    if len(bad):
        print("   bad gradf entries:", len(bad), "first", bad[:5].tolist(), "points", bad[:, 2].unique()[:10].tolist())
    for name, p in mod.named_parameters():
        k = name[len('local_aggregation_operator.'):]
        print("   grad", k, rel(p.grad.cpu(), orc.grads()[k]), float(orc.grads()[k].abs().max()))

XYZ = dict(pospool=dict(position_embedding='xyz', reduction='avg'))
SC = dict(pospool=dict(position_embedding='sin_cos', reduction='avg'))
AW = dict(adaptive_weight=dict(reduction='avg'))
case('pospool', SC, 2, 3000, 40, 144, 1)
case('pospool', XYZ, 2, 3000, 40, 144, 1)
case('pospool', SC, 2, 1024, 40, 144, 1)
case('pospool', SC, 2, 3000, 16, 144, 1)
case('pseudo_grid', {}, 2, 3000, 26, 72, 2)
case('pseudo_grid', dict(pseudo_grid=dict(KP_influence='constant')), 2, 800, 16, 36, 3)
case('pseudo_grid', {}, 2, 1500, 16, 144, 4)
case('pospool', XYZ, 3, 2400, 16, 72, 31, M=600, radius=0.15)
case('adaptive_weight', AW, 3, 2400, 16, 72, 31, M=600, radius=0.15)
case('adaptive_weight', AW, 2, 3000, 32, 72, 5)
