"""Autograd ops and groupers with the reference's names and signatures.

Mirror of /root/reference/lib/pointnet2/pointnet2_utils.py (line numbers cited per symbol): the six
``Function.apply`` aliases (furthest_point_sample, gather_operation, three_nn, three_interpolate,
grouping_operation, ball_query) plus ``QueryAndGroup`` / ``GroupAll``.  Everything runs on the
sm_100a kernels behind ``3dvlp_b200._ext``; there is no CPU path.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _ext


class FurthestPointSampling(Function):
    """pointnet2_utils.py:51-78.  xyz (B,N,3) -> (B,npoint) int32; not differentiable."""

    @staticmethod
    def forward(ctx, xyz, npoint):
        out = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class FurthestPointSampleGather(Function):
    """FPS fused with the xyz gather that follows it in every SA module
    (pointnet2_modules.py:233-240: ``gather_operation(xyz_flipped, furthest_point_sample(xyz, npoint))``)
    in one launch: xyz (B,N,3) -> (new_xyz (B,npoint,3), inds (B,npoint) int32).  Like the
    reference's composition, new_xyz is differentiable w.r.t. xyz (scatter-add of the incoming
    gradient, GatherOperation.backward :108-115) and the indices are not."""

    @staticmethod
    def forward(ctx, xyz, npoint):
        inds, new_xyz = _ext.furthest_point_sampling_gather(xyz, npoint)
        ctx.for_backwards = (inds, xyz.size(1))
        ctx.mark_non_differentiable(inds)
        return new_xyz, inds

    @staticmethod
    def backward(ctx, grad_xyz, _grad_inds=None):
        inds, n = ctx.for_backwards
        g = _ext.gather_points_grad(grad_xyz.transpose(1, 2).contiguous(), inds, n)  # (B, 3, n)
        return g.transpose(1, 2).contiguous(), None


furthest_point_sample_gather = FurthestPointSampleGather.apply


class GatherOperation(Function):
    """pointnet2_utils.py:81-115.  features (B,C,N), idx (B,npoint) -> (B,C,npoint)."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.for_backwards = (idx, features.size(1), features.size(2))
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, _, n = ctx.for_backwards
        return _ext.gather_points_grad(grad_out.contiguous(), idx, n), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """pointnet2_utils.py:118-147.  Returns (dist, idx) with dist = sqrt of the kernel's d^2."""

    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """pointnet2_utils.py:150-204.  features (B,c,m), idx/weight (B,n,3) -> (B,c,n)."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.three_interpolate_for_backward = (idx, weight, features.size(2))
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """pointnet2_utils.py:207-255.  features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.for_backwards = (idx, features.size(2))
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.for_backwards
        return _ext.group_points_grad(grad_out.contiguous(), idx, n), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """pointnet2_utils.py:258-287.  Note the argument order (radius, nsample, xyz, new_xyz)."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        out = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """Ball query + grouping (+ recentring, optional /radius, xyz channels first).

    pointnet2_utils.py:290-372, same constructor arguments and return convention: the grouped
    features, then (if requested) the recentred neighbour coordinates and the per-ball unique counts.
    """

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        if ret_unique_cnt and not sample_uniformly:
            raise AssertionError("ret_unique_cnt needs sample_uniformly")
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt

    def _resample_uniformly(self, idx):
        """Replace the first-hit padding of every ball by uniform draws from its distinct members
        (:332-341; a host loop in the reference too -- off everywhere on the hot path)."""
        counts = torch.zeros(idx.shape[:2])
        for b, ball_rows in enumerate(idx):
            for j, row in enumerate(ball_rows):
                members = torch.unique(row)
                counts[b, j] = members.numel()
                extra = torch.randint(0, members.numel(), (self.nsample - members.numel(),), dtype=torch.long)
                idx[b, j] = torch.cat((members, members[extra.to(members.device)]))
        return counts

    def forward(self, xyz, new_xyz, features=None):
        """xyz (B,N,3), new_xyz (B,npoint,3), features (B,C,N) -> (B, 3 + C, npoint, nsample)."""
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        unique_cnt = self._resample_uniformly(idx) if self.sample_uniformly else None

        local = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B, 3, npoint, nsample)
        local = local - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            local = local / self.radius

        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            grouped = local
        else:
            grouped = grouping_operation(features, idx)
            if self.use_xyz:
                grouped = torch.cat([local, grouped], dim=1)

        extras = ([local] if self.ret_grouped_xyz else []) + ([unique_cnt] if self.ret_unique_cnt else [])
        return (grouped, *extras) if extras else grouped


class GroupAll(nn.Module):
    """One group holding every point (npoint=None layers): pointnet2_utils.py:375-421.  (The
    reference forgets to store ret_grouped_xyz -- SURVEY.md appendix C-2 -- so its path raises;
    here the flag is kept.)"""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        coords = xyz.transpose(1, 2).unsqueeze(2)  # (B, 3, 1, N)
        if features is None:
            out = coords
        else:
            out = features.unsqueeze(2)
            if self.use_xyz:
                out = torch.cat([coords, out], dim=1)
        return (out, coords) if self.ret_grouped_xyz else out
