"""Pointnet2Backbone: the VoteNet-style 4 x SA + 2 x FP feature extractor.

Drop-in for /root/reference/models/base_module/backbone_module.py:11-135 -- same constructor
argument, the same hard-coded layer hyper-parameters (:29-67), the same ``data_dict`` keys in and
out, and the same parameter names (96 state-dict entries) so reference checkpoints load.
"""
import torch
import torch.nn as nn

from .pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes


class Pointnet2Backbone(nn.Module):
    def __init__(self, input_feature_dim=0):
        super().__init__()
        self.input_feature_dim = input_feature_dim
        self.sa1 = PointnetSAModuleVotes(npoint=2048, radius=0.2, nsample=64,
                                         mlp=[input_feature_dim, 64, 64, 128], use_xyz=True, normalize_xyz=True)
        self.sa2 = PointnetSAModuleVotes(npoint=1024, radius=0.4, nsample=32,
                                         mlp=[128, 128, 128, 256], use_xyz=True, normalize_xyz=True)
        self.sa3 = PointnetSAModuleVotes(npoint=512, radius=0.8, nsample=16,
                                         mlp=[256, 128, 128, 256], use_xyz=True, normalize_xyz=True)
        self.sa4 = PointnetSAModuleVotes(npoint=256, radius=1.2, nsample=16,
                                         mlp=[256, 128, 128, 256], use_xyz=True, normalize_xyz=True)
        self.fp1 = PointnetFPModule(mlp=[256 + 256, 256, 256])
        self.fp2 = PointnetFPModule(mlp=[256 + 256, 256, 256])

    def _break_up_pc(self, pc):
        xyz = pc[..., :3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def forward(self, data_dict):
        """data_dict['point_clouds']: (B, N, 3 + input_feature_dim) fp32, rows (x, y, z, features...).
        Adds sa{1,2}_{inds,xyz,features}, sa{3,4}_{xyz,features}, fp2_{features,xyz,inds}."""
        pc = data_dict["point_clouds"]
        if not self.training and not torch.is_grad_enabled() and pc.is_cuda:
            from . import fused
            out = fused.backbone_forward(self, data_dict)  # sampling chain on a side stream
            if out is not None:
                return out
        xyz, features = self._break_up_pc(pc)

        xyz, features, fps_inds = self.sa1(xyz, features)
        data_dict['sa1_inds'], data_dict['sa1_xyz'], data_dict['sa1_features'] = fps_inds, xyz, features
        xyz, features, fps_inds = self.sa2(xyz, features)
        data_dict['sa2_inds'], data_dict['sa2_xyz'], data_dict['sa2_features'] = fps_inds, xyz, features
        xyz, features, fps_inds = self.sa3(xyz, features)
        data_dict['sa3_xyz'], data_dict['sa3_features'] = xyz, features
        xyz, features, fps_inds = self.sa4(xyz, features)
        data_dict['sa4_xyz'], data_dict['sa4_features'] = xyz, features

        features = self.fp1(data_dict['sa3_xyz'], data_dict['sa4_xyz'],
                            data_dict['sa3_features'], data_dict['sa4_features'])
        features = self.fp2(data_dict['sa2_xyz'], data_dict['sa3_xyz'], data_dict['sa2_features'], features)
        data_dict['fp2_features'] = features
        data_dict['fp2_xyz'] = data_dict['sa2_xyz']
        num_seed = data_dict['fp2_xyz'].shape[1]
        # seed indices into the input cloud: SA2's FPS over SA1's output starts with SA1's own order
        data_dict['fp2_inds'] = data_dict['sa1_inds'][:, 0:num_seed]
        return data_dict
