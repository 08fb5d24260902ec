"""oracle/model_oracle.py -- torch restatement of the reference's whole networks (unfused, like the reference).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The travelling checker for the whole-model parity tests on the
GPU box, where /root/reference does not exist: the reference's ResNet backbone, classifier and scene-segmentation
head written functionally over a reference state dict, on top of oracle/la_oracle.py and a pluggable `ext`
(oracle.ext on CPU tensors, or the reference's own compiled CUDA extension from oracle/_ref on CUDA tensors).

Follows (reference, /root/reference/pytorch):
  models/backbones/resnet.py:47-68     Bottleneck.forward
  models/backbones/resnet.py:144-188   ResNet.forward
  models/heads/classifier.py:11-14,50-52   MaskedGlobalAvgPool1d / ClassifierResNet.forward
  models/heads/segmentation_head.py:61-94  SceneSegHeadResNet.forward
  models/build.py:54-57,114-117        ClassificationModel / SceneSegmentationModel.forward
Pinned against the unmodified reference models on CPU by tests/test_model_cpu.py (container).
"""
import torch
import torch.nn.functional as F

from . import la_oracle


class OracleModel:
    """state: the reference model's state_dict (keys 'backbone....', 'classifier....' / 'segmentation_head....').
    Floating-point non-buffer entries get requires_grad so .grads() mirrors the reference's parameter gradients."""

    _BUFFERS = ("running_mean", "running_var", "num_batches_tracked", "K_points")

    def __init__(self, ext, cfg, state, task, device="cpu"):
        self.ext, self.cfg, self.task = ext, cfg, task
        self.st = {}
        for k, v in state.items():
            v = v.detach().clone().to(device)
            if v.is_floating_point() and not k.endswith(self._BUFFERS):
                v.requires_grad_(True)
            self.st[k] = v
        self.training = True
        self.queries = 0   # number of ball queries issued (the reference issues 14 per backbone forward)

    # ---- building blocks ------------------------------------------------------------------------
    def _bn(self, x, prefix, momentum=0.1):
        st = self.st
        y = F.batch_norm(x, st[prefix + ".running_mean"], st[prefix + ".running_var"], st[prefix + ".weight"],
                         st[prefix + ".bias"], self.training, momentum, 1e-5)
        if self.training:
            st[prefix + ".num_batches_tracked"] += 1
        return y

    def _conv_bn(self, x, prefix, relu, momentum=None):
        """nn.Sequential(Conv1d 1x1 no bias, BatchNorm1d[, ReLU])"""
        y = F.conv1d(x, self.st[prefix + ".0.weight"])
        y = self._bn(y, prefix + ".1", self.cfg.bn_momentum if momentum is None else momentum)
        return F.relu(y) if relu else y

    def _la(self, prefix, cin, radius, nsample, q_xyz, s_xyz, q_mask, s_mask, feats):
        p = prefix + ".local_aggregation_operator."
        sub = {k[len(p):]: v for k, v in self.st.items() if k.startswith(p)}
        self.queries += 1
        return la_oracle.FAMILIES[self.cfg.local_aggregation_type](
            self.ext, sub, self.cfg, cin, cin, radius, nsample, q_xyz, s_xyz, q_mask, s_mask, feats, self.training)

    def _bottleneck(self, prefix, cin, cout, radius, nsample, xyz, mask, feats, downsample=False, sampleDl=None,
                    npoint=None):
        """resnet.py:47-68"""
        ratio = self.cfg.bottleneck_ratio
        if downsample:
            self.queries += 1
            q_xyz, q_mask, identity = la_oracle.masked_max_pool(self.ext, xyz, mask, feats, npoint, radius, nsample, sampleDl)
        else:
            q_xyz, q_mask, identity = xyz, mask, feats
        out = self._conv_bn(feats, prefix + ".conv1", relu=True)
        out = self._la(prefix + ".local_aggregation", cout // ratio, radius, nsample, q_xyz, xyz, q_mask, mask, out)
        out = self._conv_bn(out, prefix + ".conv2", relu=False)
        if cin != cout:
            identity = self._conv_bn(identity, prefix + ".shortcut", relu=False)
        return q_xyz, q_mask, F.relu(out + identity)

    def backbone(self, xyz, mask, feats):
        """resnet.py:144-188 (constructor arithmetic :97-142)"""
        c = self.cfg
        radius, sampleDl, width, ns, npts = c.radius, c.sampleDl, c.width, c.nsamples, c.npoints
        ep = {}
        f = self._conv_bn(feats, "backbone.conv1", relu=True)
        f = self._la("backbone.la1", width // 2, radius, ns[0], xyz, xyz, mask, mask, f)
        xyz, mask, f = self._bottleneck("backbone.btnk1", width // 2, width, radius, ns[0], xyz, mask, f)
        ep["res1_xyz"], ep["res1_mask"], ep["res1_features"] = xyz, mask, f
        for stage in range(4):
            sampleDl *= 2
            pre = f"backbone.layer{stage + 1}"
            xyz, mask, f = self._bottleneck(pre + ".strided_bottleneck", width, 2 * width, radius, ns[stage], xyz, mask, f,
                                            downsample=True, sampleDl=sampleDl, npoint=npts[stage])
            radius, width = radius * 2, width * 2
            for i in range(c.depth - 1):
                xyz, mask, f = self._bottleneck(pre + f".bottlneck{i}", width, width, radius, ns[stage + 1], xyz, mask, f)
            ep[f"res{stage + 2}_xyz"], ep[f"res{stage + 2}_mask"], ep[f"res{stage + 2}_features"] = xyz, mask, f
        return ep

    def classifier(self, ep):
        """classifier.py:11-14,50-52 (eval: dropout off; training: the caller must seed identically -- the tests
        compare in eval mode or with dropout p = 0 patched on both sides)"""
        st = self.st
        x = ep["res5_features"].sum(-1) / ep["res5_mask"].sum(-1)[:, None]
        for i in (0, 4, 8):
            x = F.linear(x, st[f"classifier.classifier.{i}.weight"], st[f"classifier.classifier.{i}.bias"])
            x = F.relu(self._bn(x, f"classifier.classifier.{i + 1}"))
        return F.linear(x, st["classifier.classifier.12.weight"], st["classifier.classifier.12.bias"])

    def seg_head(self, ep):
        """segmentation_head.py:61-94"""
        f = ep["res5_features"]
        for i in range(4):
            fine, coarse = f"res{4 - i}", f"res{5 - i}"
            f = la_oracle.masked_upsample_nearest(self.ext, ep[fine + "_xyz"], ep[coarse + "_xyz"], ep[fine + "_mask"],
                                                  ep[coarse + "_mask"], f)
            f = self._conv_bn(torch.cat([f, ep[fine + "_features"]], 1), f"segmentation_head.up_conv{i}", relu=True,
                              momentum=0.1)
        f = self._conv_bn(f, "segmentation_head.head", relu=True, momentum=0.1)
        return F.conv1d(f, self.st["segmentation_head.head.3.weight"], self.st["segmentation_head.head.3.bias"])

    def __call__(self, xyz, mask, feats):
        ep = self.backbone(xyz, mask, feats)
        self.end_points = ep
        return self.classifier(ep) if self.task == "classification" else self.seg_head(ep)

    def grads(self):
        return {k: v.grad for k, v in self.st.items() if v.requires_grad and v.grad is not None}
