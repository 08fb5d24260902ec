"""CPU tests (-m "not gpu", container only: they import the UNMODIFIED reference `models/`):
  * closerlook3d_b200/backbone.py has exactly the reference networks' state-dict keys and shapes (checkpoints load);
  * oracle/model_oracle.py reproduces the reference networks' forward and backward on CPU (it is the travelling
    whole-model checker of the GPU tests, so it is pinned here against the real thing)."""
import copy
import os

import pytest
import torch

from closerlook3d_b200 import synth

REF = os.path.isdir("/root/reference/pytorch/models")
pytestmark = pytest.mark.skipif(not REF, reason="/root/reference not present (GPU box)")


def _ref_cfg(ns, task, la, **over):
    from oracle import ref_loader
    cfg = ref_loader.make_config(la, **over)
    if task == "classification":
        cfg.head, cfg.num_classes, cfg.input_features_dim = "resnet_cls", 7, 3
    else:
        cfg.head, cfg.num_classes, cfg.input_features_dim = "resnet_scene_seg", 5, 4
    cfg.backbone, cfg.width, cfg.depth, cfg.bottleneck_ratio = "resnet", 12, 2, 2
    cfg.radius, cfg.sampleDl = 0.12, 0.06
    cfg.nsamples, cfg.npoints = [8, 9, 10, 9, 8], [96, 40, 16, 6]
    return cfg


def _inputs(B, N, cin, seed):
    d = synth.make_cloud_batch(B, N, cin, seed)
    return d["xyz"], d["mask"], d["features"]


@pytest.mark.parametrize("task,la,over", [
    ("classification", "pospool", dict(pospool=dict(position_embedding="xyz", reduction="avg"))),
    ("scene_segmentation", "adaptive_weight", dict(adaptive_weight=dict(weight_type="dp", num_mlps=1, reduction="avg"))),
    ("scene_segmentation", "pointwisemlp", dict(pointwisemlp=dict(feature_type="dp_fi_df", num_mlps=1, reduction="max"))),
])
def test_model_oracle_and_backbone_keys_match_reference(oracle_ext, task, la, over):
    from oracle import model_oracle, ref_loader
    ns = ref_loader.load()
    import importlib
    build = importlib.import_module("models.build")
    cfg = _ref_cfg(ns, task, la, **over)
    torch.manual_seed(5)
    ref = (build.ClassificationModel if task == "classification" else build.SceneSegmentationModel)(
        cfg, cfg.backbone, cfg.head, cfg.num_classes, cfg.input_features_dim, cfg.radius, cfg.sampleDl, cfg.nsamples,
        cfg.npoints, cfg.width, cfg.depth, cfg.bottleneck_ratio)
    ref.init_weights()
    for m in ref.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    sd = copy.deepcopy(ref.state_dict())

    # ---- (1) this package's networks carry the same keys and shapes, and load the reference checkpoint
    from closerlook3d_b200 import backbone as bb
    mine = (bb.ClassificationModel if task == "classification" else bb.SceneSegmentationModel)(cfg)
    msd = mine.state_dict()
    assert set(msd.keys()) == set(sd.keys())
    assert all(tuple(msd[k].shape) == tuple(sd[k].shape) for k in sd)
    assert list(msd.keys()) == list(sd.keys())
    mine.load_state_dict(sd, strict=True)

    # ---- (2) the oracle restatement == the reference model, forward and backward, on CPU
    xyz, mask, feats = _inputs(3, 256, cfg.input_features_dim, 12)
    ref.train()
    f_ref = feats.clone().requires_grad_(True)
    out_ref = ref(xyz, mask, f_ref)
    orc = model_oracle.OracleModel(oracle_ext, cfg, sd, task)
    f_o = feats.clone().requires_grad_(True)
    out_o = orc(xyz, mask, f_o)
    assert orc.queries == 14                                   # 10 LocalAggregation + 4 MaskedMaxPool
    assert out_o.shape == out_ref.shape
    assert float((out_o - out_ref).abs().max()) <= 1e-6 * max(1.0, float(out_ref.abs().max()))
    g = torch.randn(out_ref.shape, generator=torch.Generator().manual_seed(1))
    (out_ref * g).sum().backward()
    (out_o * g).sum().backward()
    assert float((f_o.grad - f_ref.grad).abs().max()) <= 1e-5 * max(1.0, float(f_ref.grad.abs().max()))
    og = orc.grads()
    for name, p in ref.named_parameters():
        assert name in og, name
        assert float((og[name] - p.grad).abs().max()) <= 1e-5 * max(1.0, float(p.grad.abs().max())), name
    sd_after = ref.state_dict()
    for k, v in orc.st.items():
        if k.endswith(("running_mean", "running_var")):
            assert float((v - sd_after[k]).abs().max()) <= 1e-6 * max(1.0, float(sd_after[k].abs().max())), k
