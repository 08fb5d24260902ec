"""CUDA-graph stepping (closerlook3d_b200/graphed.py): a replay must produce exactly what the eager step produces,
and the two slots of the PipelinedTrainer must each report the gradients of THEIR batch (each slot owns the
gradient tensors its graph writes; p.grad only aliases the slot captured last)."""
import numpy as np
import pytest
import torch

from closerlook3d_b200 import synth
from closerlook3d_b200.config import la_config

pytestmark = pytest.mark.gpu


def _module(cuda, la, over, C, N, K):
    from closerlook3d_b200.local_aggregation_operators import LocalAggregation
    torch.manual_seed(3)
    np.random.seed(3)
    return LocalAggregation(C, C, synth.ball_radius(N, K), K, la_config(la, **over)).to(cuda).train()


def _eager(mod, b, gout):
    for p in mod.parameters():
        p.grad = None
    f = b["features"].clone().requires_grad_(True)
    out = mod(b["xyz"], b["xyz"], b["mask"], b["mask"], f)
    out.backward(gout)
    return out.detach().clone(), f.grad.clone(), [p.grad.clone() for p in mod.parameters()]


@pytest.mark.parametrize("la,over", [
    ("pseudo_grid", dict()),
    ("adaptive_weight", dict(adaptive_weight=dict(weight_type="dp", num_mlps=1, shared_channels=1, reduction="avg"))),
    ("pointwisemlp", dict(pointwisemlp=dict(feature_type="dp_fi_df", num_mlps=1, reduction="max"))),
])
def test_pipelined_trainer_slots_report_their_own_gradients(cuda, la, over):
    from closerlook3d_b200 import pt_utils
    from closerlook3d_b200.graphed import GraphedStep, PipelinedTrainer
    B, N, K, C = 2, 2304, 16, 72
    mod = _module(cuda, la, over, C, N, K)
    batches = [{k: v.to(cuda) for k, v in synth.make_cloud_batch(B, N, C, 100 + i).items()} for i in range(3)]
    host = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in batches]
    gout = torch.randn(B, C, N, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
    was = pt_utils.cache_enabled
    pt_utils.cache_enabled = False
    try:
        ref = [_eager(mod, b, gout) for b in batches]
        gs = GraphedStep(mod, batches[0]["xyz"], batches[0]["mask"], batches[0]["features"], gout)
        seen = []
        tr = PipelinedTrainer(mod, batches[0]["xyz"], batches[0]["mask"], batches[0]["features"], gout,
                              after_step=lambda slot: seen.append(slot))
        results = []
        for i, hb in enumerate(host):
            tr.step(hb)
            torch.cuda.synchronize()
            slot = seen[-1]
            results.append((slot.out.clone(), slot.features.grad.clone(), [g.clone() for g in slot.grads()],
                            tr.result_host[i & 1].clone()))
        tr.flush()
        # the stand-alone graph over the same module still owns its gradients after the trainer captured two more
        gs.load(batches[1]["xyz"], batches[1]["mask"], batches[1]["features"])
        gs.replay()
        torch.cuda.synchronize()
        results.append((gs.out.clone(), gs.features.grad.clone(), [g.clone() for g in gs.grads()], None))
        refs = ref + [ref[1]]
        for (out, gf, gp, host_res), (o_r, gf_r, gp_r) in zip(results, refs):
            assert torch.allclose(out, o_r, rtol=0, atol=2e-6 * max(1.0, float(o_r.abs().max())))
            # backward lists are ordered by atomics -> fp32 summation order differs run to run: tolerance, not equality
            assert float((gf - gf_r).abs().max()) <= 1e-5 * max(1.0, float(gf_r.abs().max()))
            for a, b in zip(gp, gp_r):
                assert float((a - b).abs().max()) <= 5e-5 * max(1.0, float(b.abs().max()))
            if host_res is not None:   # D2H payload = [sum(out) | flat parameter gradients]
                flat = torch.cat([g.reshape(-1) for g in gp]).cpu()
                assert torch.equal(host_res[1:], flat)
                assert abs(float(host_res[0]) - float(out.sum())) <= 1e-3 * max(1.0, abs(float(out.sum())))
        # different batches must give different gradients (guards against a slot returning stale tensors)
        assert float((results[0][2][0] - results[1][2][0]).abs().max()) > 0
    finally:
        pt_utils.cache_enabled = was
