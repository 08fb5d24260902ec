"""torchrun --nproc-per-node 2 tools/test_graph_allreduce.py : NCCL all-reduce captured in a CUDA graph (GraphedStep)
against the eager all-reduce, on a tiny module.  Prints one line per rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from closerlook3d_b200 import synth, pt_utils
from closerlook3d_b200.config import la_config
from closerlook3d_b200.local_aggregation_operators import LocalAggregation
from closerlook3d_b200.graphed import GraphedStep
torch.manual_seed(0); np.random.seed(0)
B, N, K, C = 2, 2304, 16, 72
mod = LocalAggregation(C, C, synth.ball_radius(N, K), K, la_config("adaptive_weight", adaptive_weight=dict(weight_type="dp", num_mlps=1, shared_channels=1, reduction="avg"))).to(dev).train()
d = {k: v.to(dev) for k, v in synth.make_cloud_batch(B, N, C, 10 + rank, b_offset=rank * B).items()}
gout = torch.randn(B, C, N, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
pt_utils.cache_enabled = False
# eager reference: local grads, then all-reduce (sum / world)
for p in mod.parameters(): p.grad = None
f = d["features"].clone().requires_grad_(True)
mod(d["xyz"], d["xyz"], d["mask"], d["mask"], f).backward(gout)
ref = torch.cat([p.grad.reshape(-1) for p in mod.parameters()]).clone()
dist.all_reduce(ref); ref /= world
print(f"rank {rank}: eager all-reduce ok", flush=True)
gs = GraphedStep(mod, d["xyz"], d["mask"], d["features"], gout)
print(f"rank {rank}: captured (allreduce in graph = {gs.allreduce})", flush=True)
for _ in range(3):
    gs.replay()
torch.cuda.synchronize()
err = float((gs.flat_grad - ref).abs().max()) / max(1e-12, float(ref.abs().max()))
print(f"rank {rank}: graph all-reduce vs eager: rel err {err:.2e}", flush=True)
assert err < 1e-4
del gs  # a live graph with a captured collective keeps ncclCommDestroy waiting
import gc; gc.collect(); torch.cuda.synchronize()
dist.barrier(); dist.destroy_process_group()
print(f'rank {rank}: clean exit', flush=True)
