"""Host side of the fused PointWiseMLP (feature_type 'dp_fi_df', num_mlps 1, reduction 'max').

Reference: /root/reference/pytorch/models/local_aggregation_operators.py:254-257,288-303.
The conv weight W (Cout, 3+2C) = [Wp | Wc | Wr] and sgn = sign(BN gamma) are folded once per call into
    wcat (2*Cop, C+3):  rows 0..Cout-1       = [Wc - Wr | 0 0 0]        -> A  = (Wc-Wr) f
                        rows Cop..Cop+Cout-1 = sgn * [Wr | Wp]          -> T  = sgn (Wr f + Wp s/r)
so that ONE per-point product over the augmented point-major matrix [f | s/r] yields both terms
(csrc/pwmlp.cu explains why y[o,q,k] = a'[q][o] + sgn*T[j_k][o] is the reference's function).  Only this
folding / unfolding of the small weight matrices is done with torch ops; all per-point and per-neighbour work
is in libcl3d.
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
        sgn = torch.where(bn_weight >= 0, 1.0, -1.0).to(torch.float32).contiguous()
        wcat = torch.zeros(2 * Cop, C + 3, dtype=torch.float32, device=features.device)
        wcat[:Cout, :C] = W[:, 3:3 + C] - W[:, 3 + C:]
        wcat[Cop:Cop + Cout, :C] = sgn[:, None] * W[:, 3 + C:]
        wcat[Cop:Cop + Cout, C:] = sgn[:, None] * wp
        fa_pm = ops.to_point_major_aug(features, support_xyz, radius)          # (B,N,Cpa)
        Cpa = fa_pm.shape[2]
        # AB[p][n] = sum_c fa[p][c] * wcat[n][c]
        ab_pm = ops.sgemm(fa_pm, Cpa, 1, wcat, 1, C + 3, B * N, 2 * Cop, C + 3).view(B, N, 2 * Cop)
        ysel, aq, sq, karg, partial = ops.pwmlp_fwd_stats(ab_pm, wp, sgn, query_xyz, nl.idx, Cout, radius)
        training = bn.training or (bn.running_mean is None)
        momentum = bn.momentum if bn.momentum is not None else 0.0
        if training and bn.running_mean is not None:
            bn.num_batches_tracked.add_(1)
            if bn.momentum is None:
                momentum = 1.0 / float(bn.num_batches_tracked)
        stats = ops.bn_finalize(partial, Cout, B * M * K, bn.eps, momentum, training, bn.running_mean,
                                bn.running_var)
        out = ops.pwmlp_fwd_out(ysel, stats, bn_weight, bn_bias)
        ctx.nl, ctx.radius, ctx.training, ctx.dims = nl, radius, training, (B, C, N, M, K, Cout, Cop, Cpa)
        ctx.save_for_backward(out, fa_pm, ab_pm, wp, sgn, wcat, ysel, aq, sq, karg, stats, bn_weight, query_xyz)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        out, fa_pm, ab_pm, wp, sgn, wcat, ysel, aq, sq, karg, stats, bn_weight, query_xyz = ctx.saved_tensors
        B, C, N, M, K, Cout, Cop, Cpa = ctx.dims
        if not ctx.training:
            raise NotImplementedError("PointWiseMLP backward in eval mode (running statistics) is not fused")
        nl = ctx.nl
        off, ent = nl.csr_all_slots()
        grad_ab, grad_wp, dgamma, dbeta = ops.pwmlp_bwd(grad_out.contiguous(), out, ab_pm, wp, sgn, query_xyz, nl.idx,
                                                        off, ent, ysel, aq, sq, karg, stats, bn_weight, ctx.radius)
        P = B * N
        Cp = ops.padded_channels(C)
        # d/dfeat (point-major) = grad_AB (P x 2Cop) @ wcat[:, :C] (2Cop x C)
        gf_pm = ops.sgemm(grad_ab, 2 * Cop, 1, wcat, C + 3, 1, P, C, 2 * Cop, ldc=Cp).view(B, N, Cp)
        grad_feat = ops.to_channel_major(gf_pm, C)
        # d/dwcat (2Cop x (C+3)) = grad_AB^T (2Cop x P) @ fa (P x (C+3)): long reduction -> split-K
        splitk = max(1, min(256, P // 256))
        gwcat = ops.sgemm(grad_ab, 1, 2 * Cop, fa_pm, Cpa, 1, 2 * Cop, C + 3, P, splitk=splitk)
        gA = gwcat[:Cout, :C]
        gT = sgn[:, None] * gwcat[Cop:Cop + Cout]
        gW = torch.cat([gT[:, C:] + grad_wp.t(), gA, gT[:, :C] - gA], dim=1).view(Cout, 3 + 2 * C, 1, 1)
        return grad_feat, gW, dgamma, dbeta, None, None, None, None, None


def forward(module, query_xyz, support_xyz, query_mask, support_mask, support_features):
    nl = pt_utils.neighbors(query_xyz, support_xyz, query_mask, support_mask, module.radius, module.nsample)
    conv, bn = module.mlps.conv0[0], module.mlps.conv0[1]
    return _FusedPointWiseMLP.apply(support_features.contiguous(), conv.weight, bn.weight, bn.bias, nl, query_xyz,
                                    support_xyz, module.radius, bn)
