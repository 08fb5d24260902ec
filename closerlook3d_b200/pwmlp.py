"""Host side of the fused PointWiseMLP (feature_type 'dp_fi_df', num_mlps 1, reduction 'max').

Reference: /root/reference/pytorch/models/local_aggregation_operators.py:254-257,288-303.
The conv weight W (Cout, 3+2C) = [Wp | Wc | Wr] is split once per call into
    Wp   (Cout,3)                     -> added per neighbour inside the gather kernel
    Wcat (2*Cop, C) = [Wc-Wr ; Wr]    -> per-POINT products  AB = f Wcat^T  (csrc/gemm.cu)
(see csrc/pwmlp.cu for why this is the same function).  Only the split / re-assembly of the small weight
matrices is done with torch ops; all per-point and per-neighbour work is in libcl3d.
"""
import torch
from torch.autograd import Function

from . import ops, pt_utils


class _FusedPointWiseMLP(Function):
    @staticmethod
    def forward(ctx, features, conv_weight, bn_weight, bn_bias, nl, query_xyz, support_xyz, radius, bn):
        B, C, N = features.shape
        M, K = nl.idx.shape[1], nl.idx.shape[2]
        Cout = conv_weight.shape[0]
        Cop = ops.padded_channels(Cout)
        W = conv_weight.view(Cout, 3 + 2 * C)
        wp = W[:, :3].contiguous()
        wcat = torch.zeros(2 * Cop, C, dtype=torch.float32, device=features.device)
        wcat[:Cout] = W[:, 3:3 + C] - W[:, 3 + C:]
        wcat[Cop:Cop + Cout] = W[:, 3 + C:]
        feat_pm = ops.to_point_major(features)                      # (B,N,Cp)
        Cp = feat_pm.shape[2]
        # AB[p][n] = sum_c feat[p][c] * wcat[n][c]
        ab_pm = ops.sgemm(feat_pm, Cp, 1, wcat, 1, C, B * N, 2 * Cop, C).view(B, N, 2 * Cop)
        ymax, ymin, arg, partial = ops.pwmlp_fwd_stats(ab_pm, wp, query_xyz, support_xyz, nl.idx, Cout, radius)
        training = bn.training or (bn.running_mean is None)
        momentum = bn.momentum if bn.momentum is not None else 0.0
        if training and bn.running_mean is not None:
            bn.num_batches_tracked.add_(1)
            if bn.momentum is None:
                momentum = 1.0 / float(bn.num_batches_tracked)
        stats = ops.bn_finalize(partial, Cout, B * M * K, bn.eps, momentum, training, bn.running_mean,
                                bn.running_var)
        out = ops.pwmlp_fwd_out(ymax, ymin, stats, bn_weight, bn_bias)
        ctx.nl, ctx.radius, ctx.training, ctx.dims = nl, radius, training, (B, C, N, M, K, Cout, Cop, Cp)
        ctx.save_for_backward(out, feat_pm, ab_pm, wp, wcat, ymax, ymin, arg, stats, bn_weight, query_xyz,
                              support_xyz)
        ctx.mark_non_differentiable()
        return out

    @staticmethod
    def backward(ctx, grad_out):
        out, feat_pm, ab_pm, wp, wcat, ymax, ymin, arg, stats, bn_weight, query_xyz, support_xyz = ctx.saved_tensors
        B, C, N, M, K, Cout, Cop, Cp = ctx.dims
        if not ctx.training:
            raise NotImplementedError("PointWiseMLP backward in eval mode (running statistics) is not fused")
        grad_ab, grad_wp, dgamma, dbeta = ops.pwmlp_bwd(grad_out.contiguous(), out, ab_pm, wp, query_xyz,
                                                        support_xyz, ctx.nl.idx, ymax, ymin, arg, stats, bn_weight,
                                                        ctx.radius)
        P = B * N
        # d/dfeat (point-major) = grad_AB (P x 2Cop) @ wcat (2Cop x C)
        gf_pm = ops.sgemm(grad_ab, 2 * Cop, 1, wcat, C, 1, P, C, 2 * Cop, ldc=Cp).view(B, N, Cp)
        grad_feat = ops.to_channel_major(gf_pm, C)
        # d/dwcat (2Cop x C) = grad_AB^T (2Cop x P) @ feat (P x C): long reduction -> split-K
        splitk = max(1, min(256, P // 256))
        gwcat = ops.sgemm(grad_ab, 1, 2 * Cop, feat_pm, Cp, 1, 2 * Cop, C, P, splitk=splitk)
        gA, gB = gwcat[:Cout], gwcat[Cop:Cop + Cout]
        gW = torch.cat([grad_wp.t(), gA, gB - gA], dim=1).view(Cout, 3 + 2 * C, 1, 1)
        return grad_feat, gW, dgamma, dbeta, None, None, None, None, None


def forward(module, query_xyz, support_xyz, query_mask, support_mask, support_features):
    nl = pt_utils.neighbors(query_xyz, support_xyz, query_mask, support_mask, module.radius, module.nsample)
    conv, bn = module.mlps.conv0[0], module.mlps.conv0[1]
    return _FusedPointWiseMLP.apply(support_features.contiguous(), conv.weight, bn.weight, bn.bias, nl, query_xyz,
                                    support_xyz, module.radius, bn)
