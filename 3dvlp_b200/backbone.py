"""Pointnet2Backbone: the VoteNet-style 4 x SA + 2 x FP feature extractor.

Drop-in for /root/reference/models/base_module/backbone_module.py:11-135 -- same constructor
argument, the same layer hyper-parameters (:29-67, kept in one table below), the same ``data_dict`` keys
in and out, and the same parameter names (``sa1..sa4``, ``fp1``, ``fp2``: 96 state-dict entries) so
reference checkpoints load.
"""
import torch
import torch.nn as nn

from .pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes

# (name, npoint, radius, nsample, hidden widths) -- backbone_module.py:29-63; the first MLP width is the
# previous layer's output (input_feature_dim for sa1)
SA_LAYERS = (
    ("sa1", 2048, 0.2, 64, (64, 64, 128)),
    ("sa2", 1024, 0.4, 32, (128, 128, 256)),
    ("sa3", 512, 0.8, 16, (128, 128, 256)),
    ("sa4", 256, 1.2, 16, (128, 128, 256)),
)
FP_LAYERS = (("fp1", (256 + 256, 256, 256)), ("fp2", (256 + 256, 256, 256)))  # :65-67


class Pointnet2Backbone(nn.Module):
    def __init__(self, input_feature_dim=0):
        super().__init__()
        self.input_feature_dim = input_feature_dim
        width = input_feature_dim
        for name, npoint, radius, nsample, hidden in SA_LAYERS:
            setattr(self, name, PointnetSAModuleVotes(npoint=npoint, radius=radius, nsample=nsample,
                                                      mlp=[width, *hidden], use_xyz=True, normalize_xyz=True))
            width = hidden[-1]
        for name, spec in FP_LAYERS:
            setattr(self, name, PointnetFPModule(mlp=list(spec)))

    def _break_up_pc(self, pc):
        """(B, N, 3 + C) rows -> xyz (B, N, 3) and channel-major features (B, C, N) (None when C = 0)."""
        feats = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return pc[..., :3].contiguous(), feats

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
        for name, *_ in SA_LAYERS:
            xyz, features, fps_inds = getattr(self, name)(xyz, features)
            data_dict[name + "_xyz"] = xyz
            data_dict[name + "_features"] = features
            if name in ("sa1", "sa2"):  # the reference keeps the indices of the first two layers only
                data_dict[name + "_inds"] = fps_inds

        up = self.fp1(data_dict["sa3_xyz"], data_dict["sa4_xyz"], data_dict["sa3_features"], data_dict["sa4_features"])
        up = self.fp2(data_dict["sa2_xyz"], data_dict["sa3_xyz"], data_dict["sa2_features"], up)
        data_dict["fp2_features"] = up
        data_dict["fp2_xyz"] = data_dict["sa2_xyz"]
        # seed indices into the input cloud: SA2's FPS over SA1's output starts with SA1's own order
        # (backbone_module.py:108,134)
        data_dict["fp2_inds"] = data_dict["sa1_inds"][:, : data_dict["fp2_xyz"].shape[1]].contiguous()
        return data_dict
