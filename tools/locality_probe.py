"""How much would spatially ordered execution buy?  Times one LocalAggregation step (fwd+bwd, eager, per entry
point) on a BASELINE config with the clouds as generated (random point order) and with every cloud's points
pre-sorted along a Morton curve (same data, permuted consistently) -- the second is the upper bound for a
kernel-side spatial execution order (neighbour rows hit L1 instead of L2)."""
import sys, json
import torch
sys.path.insert(0, ".")
from closerlook3d_b200 import synth, _lib, pt_utils
import bench


def morton_perm(xyz):
    q = (xyz.clamp(0, 0.999999) * 1024).long()
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    key = spread(q[..., 0]) | (spread(q[..., 1]) << 1) | (spread(q[..., 2]) << 2)
    return key.argsort(dim=1, stable=True)


def run(ci, sort):
    d = synth.baseline_inputs(ci)
    spec = d["spec"]
    xyz, mask, feats = d["xyz"], d["mask"], d["features"]
    if sort:
        # keep the valid-prefix structure: sort only inside the valid prefix of every cloud
        for b in range(xyz.shape[0]):
            nv = int(mask[b].sum())
            p = morton_perm(xyz[b:b + 1, :nv])[0]
            xyz[b, :nv] = xyz[b, :nv][p]
            feats[b, :, :nv] = feats[b][:, :nv][:, p]
    dev = torch.device("cuda:0")
    mod, _ = bench.build_module(spec, dev)
    pt_utils.cache_enabled = False
    xyz, mask = xyz.to(dev), mask.to(dev)
    f = feats.to(dev).requires_grad_(True)
    def step():
        f.grad = None
        out = mod(xyz, xyz, mask, mask, f)
        out.sum().backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    _lib.profiler.start()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    prof = _lib.profiler.stop()
    return {k: round(t / 5, 4) for k, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])}


if __name__ == "__main__":
    for ci in [int(a) for a in sys.argv[1:]] or [2]:
        for sort in (False, True):
            print("config", ci, "sorted" if sort else "random", json.dumps(run(ci, sort)))
