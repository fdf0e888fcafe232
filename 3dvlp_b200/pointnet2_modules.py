"""Set-abstraction and feature-propagation modules with the reference's API.

Mirror of /root/reference/lib/pointnet2/pointnet2_modules.py: ``PointnetSAModuleVotes`` (:164-272),
``PointnetFPModule`` (:356-416) and the variants built from the same ops (``PointnetSAModuleMSG``
:78-124, ``PointnetSAModule`` :127-161, ``PointnetSAModuleMSGVotes`` :274-353,
``PointnetLFPModuleMSG`` :418-496).  Same constructor keywords, forward signatures, return tuples
and parameter names, so checkpoints and call sites carry over unchanged.

Two execution paths share the parameters:
  * the *generic* path composes the individual ops (autograd-capable: training, odd poolings);
  * the *fused* path (eval mode, max pooling; see fused.py) folds BatchNorm into the convolutions and
    runs FPS+gather, ball query and the grouped shared MLP + max-pool as sm_100a kernels.
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import mlp_train
from . import pointnet2_utils
from . import pytorch_utils as pt_utils


def _sample_centres(xyz, npoint, inds):
    """FPS (unless `inds` is given) + the xyz gather of pointnet2_modules.py:233-240."""
    if npoint is None:
        return None, inds
    if inds is None:
        # one launch; differentiable w.r.t. xyz like the reference's gather_operation (votes carry grad:
        # proposal_module_fcos.py:62-67 samples vote_xyz and decodes box centres from the result)
        new_xyz, inds = pointnet2_utils.furthest_point_sample_gather(xyz.contiguous(), npoint)
        return new_xyz, inds
    xyz_flipped = xyz.transpose(1, 2).contiguous()
    new_xyz = pointnet2_utils.gather_operation(xyz_flipped, inds).transpose(1, 2).contiguous()
    return new_xyz, inds


def _pool_max(x):
    return F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)


def _mlp_pool_max(mlp, grouped):
    """``max_pool2d(mlp(grouped))`` over the nsample axis -> (B, C, npoint): pointnet2_modules.py:251-257.
    Training mode and every call under autograd run on the tcgen05 kernels of mlp_train.py (batch-statistics
    BatchNorm, backward, weight gradients); shapes it does not cover take torch's own layers."""
    if mlp_train.supported(mlp, grouped):
        return mlp_train.shared_mlp_maxpool(mlp, grouped)
    return _pool_max(mlp(grouped))


class _PointnetSAModuleBase(nn.Module):
    """pointnet2_modules.py:26-75: returns (new_xyz, concatenated per-scale features)."""

    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def forward(self, xyz, features=None):
        new_xyz, _ = _sample_centres(xyz, self.npoint, None)
        outs = [_mlp_pool_max(mlp, grouper(xyz, new_xyz, features)) for grouper, mlp in zip(self.groupers, self.mlps)]
        return new_xyz, torch.cat(outs, dim=1)


def _build_scales(module, npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly):
    assert len(radii) == len(nsamples) == len(mlps)
    module.npoint = npoint
    module.groupers = nn.ModuleList()
    module.mlps = nn.ModuleList()
    for radius, nsample, spec in zip(radii, nsamples, mlps):
        module.groupers.append(
            pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz, sample_uniformly=sample_uniformly)
            if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
        if use_xyz:
            spec[0] += 3  # the reference mutates the caller's list too (appendix C-4)
        module.mlps.append(pt_utils.SharedMLP(spec, bn=bn))


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]],
                 bn: bool = True, use_xyz: bool = True, sample_uniformly: bool = False):
        super().__init__()
        _build_scales(self, npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly)


class PointnetSAModule(PointnetSAModuleMSG):
    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz)


class PointnetSAModuleVotes(nn.Module):
    """SA layer that also returns the sampled indices (VoteNet).  pointnet2_modules.py:164-272."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pooling: str = 'max', sigma: float = None,
                 normalize_xyz: bool = False, sample_uniformly: bool = False, ret_unique_cnt: bool = False):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.pooling = pooling
        self.use_xyz = use_xyz
        self.sigma = sigma if sigma is not None else (self.radius / 2 if self.radius is not None else None)
        self.normalize_xyz = normalize_xyz
        self.ret_unique_cnt = ret_unique_cnt
        self.bn = bn
        if npoint is not None:
            self.grouper = pointnet2_utils.QueryAndGroup(
                radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True, normalize_xyz=normalize_xyz,
                sample_uniformly=sample_uniformly, ret_unique_cnt=ret_unique_cnt)
        else:
            self.grouper = pointnet2_utils.GroupAll(use_xyz, ret_grouped_xyz=True)
        mlp_spec = mlp
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3  # in place, like the reference (:204-206)
        self.mlp_module = pt_utils.SharedMLP(mlp_spec, bn=bn)
        self._fused = None  # lazily built by fused.py

    def _can_fuse(self, xyz, features):
        return (not self.training and not torch.is_grad_enabled() and self.pooling == 'max'
                and self.npoint is not None and not self.ret_unique_cnt
                and not self.grouper.sample_uniformly and xyz.is_cuda)

    def forward(self, xyz, features=None, inds=None):
        if inds is not None:
            assert inds.shape[1] == self.npoint
        if self._can_fuse(xyz, features):
            from . import fused
            out = fused.sa_forward(self, xyz, features, inds)
            if out is not None:
                return out
        new_xyz, inds = _sample_centres(xyz, self.npoint, inds)
        if not self.ret_unique_cnt:
            grouped_features, grouped_xyz = self.grouper(xyz, new_xyz, features)
        else:
            grouped_features, grouped_xyz, unique_cnt = self.grouper(xyz, new_xyz, features)
        if self.pooling == 'max' and mlp_train.supported(self.mlp_module, grouped_features):
            # training / autograd: shared MLP with batch-statistics BatchNorm + max-pool on the tcgen05 kernels
            new_features = mlp_train.shared_mlp_maxpool(self.mlp_module, grouped_features)
            if not self.ret_unique_cnt:
                return new_xyz, new_features, inds
            return new_xyz, new_features, inds, unique_cnt
        new_features = self.mlp_module(grouped_features)  # (B, mlp[-1], npoint, nsample)
        if self.pooling == 'max':
            new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == 'avg':
            new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == 'rbf':
            # radial-basis weighting of the neighbours, normalised by nsample (:258-266)
            rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (self.sigma ** 2) / 2)
            new_features = torch.sum(new_features * rbf.unsqueeze(1), -1, keepdim=True) / float(self.nsample)
        new_features = new_features.squeeze(-1)
        if not self.ret_unique_cnt:
            return new_xyz, new_features, inds
        return new_xyz, new_features, inds, unique_cnt


class PointnetSAModuleMSGVotes(nn.Module):
    """pointnet2_modules.py:274-353."""

    def __init__(self, *, mlps: List[List[int]], npoint: int, radii: List[float], nsamples: List[int],
                 bn: bool = True, use_xyz: bool = True, sample_uniformly: bool = False):
        super().__init__()
        _build_scales(self, npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly)

    def forward(self, xyz, features=None, inds=None):
        new_xyz, inds = _sample_centres(xyz, self.npoint, inds)
        outs = [_mlp_pool_max(mlp, grouper(xyz, new_xyz, features)) for grouper, mlp in zip(self.groupers, self.mlps)]
        return new_xyz, torch.cat(outs, dim=1), inds


class PointnetFPModule(nn.Module):
    """Feature propagation: 3-NN inverse-distance interpolation + skip concat + shared MLP.
    pointnet2_modules.py:356-416."""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)
        self._fused = None

    def forward(self, unknown, known, unknow_feats, known_feats):
        if (not self.training and not torch.is_grad_enabled() and known is not None
                and unknow_feats is not None and unknown.is_cuda):
            from . import fused
            out = fused.fp_forward(self, unknown, known, unknow_feats, known_feats)
            if out is not None:
                return out
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated_feats = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated_feats = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        if unknow_feats is not None:
            new_features = torch.cat([interpolated_feats, unknow_feats], dim=1)  # (B, C2 + C1, n)
        else:
            new_features = interpolated_feats
        new_features = new_features.unsqueeze(-1)   # (B, C, n, 1): "max-pool" over one sample
        if mlp_train.supported(self.mlp, new_features):
            return mlp_train.shared_mlp_maxpool(self.mlp, new_features)
        return self.mlp(new_features).squeeze(-1)


class PointnetLFPModuleMSG(nn.Module):
    """Learnable feature propagation (pointnet2_modules.py:418-496)."""

    def __init__(self, *, mlps: List[List[int]], radii: List[float], nsamples: List[int], post_mlp: List[int],
                 bn: bool = True, use_xyz: bool = True, sample_uniformly: bool = False):
        super().__init__()
        self.post_mlp = pt_utils.SharedMLP(post_mlp, bn=bn)
        _build_scales(self, 0, radii, nsamples, mlps, bn, use_xyz, sample_uniformly)
        self.npoint = None

    def forward(self, xyz2, xyz1, features2, features1):
        outs = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            f = _mlp_pool_max(mlp, grouper(xyz1, xyz2, features1))  # (B, mlp[-1], N2)
            if features2 is not None:
                f = torch.cat([f, features2], dim=1)
            outs.append(self.post_mlp(f.unsqueeze(-1)))
        return torch.cat(outs, dim=1).squeeze(-1)
