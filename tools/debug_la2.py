import copy, sys, torch, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from closerlook3d_b200 import synth
from closerlook3d_b200.config import la_config
from closerlook3d_b200.local_aggregation_operators import LocalAggregation
import oracle; oracle.build()
from oracle import ext as oext, la_oracle
from test_local_aggregation_gpu import _randomize_bn, SINCOS_AVG
B,N,K,C = 2,3000,40,144; seed = 2000+N+C
cfg = la_config('pospool', **SINCOS_AVG)
torch.manual_seed(seed); np.random.seed(seed)
r = synth.ball_radius(N,K)
mod = LocalAggregation(C,C,r,K,cfg); _randomize_bn(mod, seed); sd = copy.deepcopy(mod.state_dict())
d = synth.make_cloud_batch(B,N,C,seed); xyz,mask,feats = d['xyz'],d['mask'],d['features']
gout = torch.randn(B,C,N, generator=torch.Generator().manual_seed(seed+2))
orc = la_oracle.OracleLocalAggregation(oext,'pospool',C,C,r,K,cfg,sd)
f_ref = feats.clone().requires_grad_(True); o_ref = orc(xyz,xyz,mask,mask,f_ref); (o_ref*gout).sum().backward()
dev = torch.device('cuda:0'); mod = mod.to(dev); f = feats.to(dev).requires_grad_(True)
out = mod(xyz.to(dev),xyz.to(dev),mask.to(dev),mask.to(dev),f); (out*gout.to(dev)).sum().backward()
o = out.detach().cpu(); 
flip = ((o>0) != (o_ref>0)).nonzero()
print("relu flips:", flip.tolist()[:10], "count", len(flip))
for b,c,q in flip.tolist()[:5]:
    print("  ", b,c,q, "mine", float(o[b,c,q]), "ref", float(o_ref[b,c,q]), "gout", float(gout[b,c,q]), "gamma", float(sd['local_aggregation_operator.out_transform.0.weight'][c]))
dg = (f.grad.cpu()-f_ref.grad).abs(); print("max gradf err", float(dg.max()), "max grad", float(f_ref.grad.abs().max()))
bad = (dg > 1e-5*f_ref.grad.abs().max()).nonzero(); print("bad count", len(bad), "channels", bad[:,1].unique().tolist()[:10], "batches", bad[:,0].unique().tolist())
