"""Proposal generation of the VoteNet-style detector: vote aggregation + ROI head + box decode.

Drop-in for /root/reference/models/proposal_module/proposal_module_fcos.py:21-178 (``ProposalModule``:
same constructor, sub-module / parameter names -- ``vote_aggregation.*``, ``proposal.convs.{0,1,3,4}``,
``proposal.{objectness,box,sem_cls,heading_cls,heading_reg}_predictor`` -- forward signature and
data_dict keys) and for models/proposal_module/ROI_heads/roi_heads.py:16-147 (``StandardROIHeads``).
SURVEY.md section 8(f) rank 1: it completes BASELINE config 2, "VoteNet detector backbone +
vote/proposal sampling".

Eval fast path on the GPU: the aggregated vote features arrive from the fused SA kernel together with
their bf16 point-major twin, the two Conv1d+BN+ReLU layers and ALL predictor heads (one concatenated
weight) are three launches of the tcgen05 linear kernel (pn2_linear_bf16; bf16 operands, fp32
accumulate), and the box decode stays on the device (box_util.decode_pred_box) where the reference
goes through ``.cpu().numpy()`` every forward (proposal_module_fcos.py:127-131).
"""
import math

import torch
import torch.nn as nn

from . import tensorcore as tcm
from .box_util import decode_pred_box
from .pointnet2_modules import PointnetSAModuleVotes
from .voting import VotingModule, _fold_conv_bn
from .backbone import Pointnet2Backbone


class StandardROIHeads(nn.Module):
    """roi_heads.py:16-147.  ``back_trace_type`` other than None raises NotImplementedError there too
    (:79-80), so the ROI-grid pooler is not part of the surface."""

    def __init__(self, num_heading_bin, num_class, use_exp=True, use_centerness=False, obj_loss=True,
                 back_trace_type=None, seed_feat_dim=256, density=1, rep_type="ray", revisit_method="one_step",
                 interp_num=16, one_step_type="maxpool-coord", use_kl_loss=False):
        super().__init__()
        self.num_heading_bin = num_heading_bin
        self.num_class = num_class
        self.use_exp = use_exp
        self.use_centerness = use_centerness
        self.obj_loss = obj_loss
        self.seed_feat_dim = seed_feat_dim
        self.back_trace_type = back_trace_type
        self.use_kl_loss = use_kl_loss
        convs = [nn.Conv1d(128, 128, kernel_size=1), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
                 nn.Conv1d(128, 128, kernel_size=1), nn.BatchNorm1d(128), nn.ReLU(inplace=True)]
        self.convs = nn.Sequential(*convs)
        self.objectness_predictor = nn.Conv1d(128, 2, kernel_size=1)
        if self.use_kl_loss:
            self.alpha_predictor = nn.Conv1d(128, 6, kernel_size=1)
            self.alpha_activation = nn.Sigmoid()
        self.box_predictor = nn.Conv1d(128, 6, kernel_size=1)
        if self.num_class:
            self.sem_cls_predictor = nn.Conv1d(128, num_class, kernel_size=1)
        self.heading_cls_predictor = nn.Conv1d(128, num_heading_bin, kernel_size=1)
        self.heading_reg_predictor = nn.Conv1d(128, num_heading_bin, kernel_size=1)
        for layer in convs:  # roi_heads.py:66-70
            if isinstance(layer, nn.Conv1d):
                nn.init.kaiming_normal_(layer.weight, mode="fan_out", nonlinearity="relu")
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, 0)
        for predictor in (self.objectness_predictor, self.box_predictor):  # :72-75
            nn.init.normal_(predictor.weight, std=0.001)
            if predictor.bias is not None:
                nn.init.constant_(predictor.bias, 0)
        if self.back_trace_type is not None:
            raise NotImplementedError()
        self._fused = None

    def _heads(self):
        """(name, module) of every predictor in the column order of the concatenated head weight."""
        heads = [("objectness", self.objectness_predictor)]
        if self.use_kl_loss:
            heads.append(("alpha", self.alpha_predictor))
        heads.append(("box", self.box_predictor))
        if self.num_class:
            heads.append(("sem_cls", self.sem_cls_predictor))
        heads += [("heading_cls", self.heading_cls_predictor), ("heading_reg", self.heading_reg_predictor)]
        return heads

    def _plan(self):
        key = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if self._fused is not None and self._fused["key"] == key:
            return self._fused
        w1, b1 = _fold_conv_bn(self.convs[0], self.convs[1])
        w2, b2 = _fold_conv_bn(self.convs[3], self.convs[4])
        ws, bs, cols, at = [], [], {}, 0
        for name, conv in self._heads():
            w, b = _fold_conv_bn(conv, None)
            cols[name] = (at, at + w.shape[0])
            at += w.shape[0]
            ws.append(w)
            bs.append(b)
        n = max(64, (at + 15) // 16 * 16)  # the linear kernel's N granularity; one 64-wide panel at least
        wh = torch.zeros(n, 128, device=w1.device)
        bh = torch.zeros(n, device=w1.device)
        wh[:at] = torch.cat(ws, 0)
        bh[:at] = torch.cat(bs, 0)
        self._fused = {"key": key, "l1": (tcm.PackedWeight(w1), b1), "l2": (tcm.PackedWeight(w2), b2),
                       "heads": (tcm.PackedWeight(wh), bh), "cols": cols}
        return self._fused

    def forward(self, ROI_features, data_dict):
        """ROI_features (B, 128, num_proposal) -> data_dict with heading_scores,
        heading_residuals(_normalized), rois, objectness_scores, bbox_mask, sem_cls_scores[, alpha]."""
        batch_size, _, num_proposal = ROI_features.shape
        if not self.training and not torch.is_grad_enabled() and ROI_features.is_cuda:
            p = self._plan()
            from . import fused
            rows = fused.cached_rows(ROI_features)
            if rows is None or rows.shape != (batch_size * num_proposal, 128):
                rows = tcm.pack_rows(ROI_features.contiguous().float(), 128)
            h = tcm.linear(rows, p["l1"][0], p["l1"][1], relu=True, out="bf16")
            h = tcm.linear(h, p["l2"][0], p["l2"][1], relu=True, out="bf16")
            y = tcm.linear(h, p["heads"][0], p["heads"][1], relu=False, out="f32").view(batch_size, num_proposal, -1)
            out = {name: y[:, :, a:b] for name, (a, b) in p["cols"].items()}
        else:  # the reference arithmetic (roi_heads.py:107-129)
            x = self.convs(ROI_features)
            out = {name: conv(x).permute(0, 2, 1) for name, conv in self._heads()}
        if self.use_kl_loss:
            data_dict["alpha"] = self.alpha_activation(out["alpha"]) * 0.1 - 0.05
        pred_box_reg = out["box"].exp()
        if self.num_class:
            data_dict["sem_cls_scores"] = out["sem_cls"]
        data_dict["heading_scores"] = out["heading_cls"]
        data_dict["heading_residuals_normalized"] = out["heading_reg"]
        data_dict["heading_residuals"] = out["heading_reg"] * (math.pi / self.num_heading_bin)
        data_dict["rois"] = pred_box_reg
        data_dict["objectness_scores"] = out["objectness"]
        data_dict["bbox_mask"] = data_dict["objectness_scores"].argmax(-1)  # 0: invalid 1: valid
        return data_dict


class ProposalModule(nn.Module):
    """proposal_module_fcos.py:21-178."""

    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr, num_proposal, sampling,
                 seed_feat_dim=256, mask_box=False, use_kl_loss=False, use_vote_weight=False):
        super().__init__()
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.mean_size_arr = mean_size_arr
        self.num_proposal = num_proposal
        self.sampling = sampling
        self.seed_feat_dim = seed_feat_dim
        self.mask_box = mask_box
        self.use_kl_loss = use_kl_loss
        self.use_vote_weight = use_vote_weight
        self.vote_aggregation = PointnetSAModuleVotes(npoint=self.num_proposal, radius=0.3, nsample=16,
                                                      mlp=[self.seed_feat_dim, 128, 128, 128], use_xyz=True,
                                                      normalize_xyz=True)
        self.proposal = StandardROIHeads(num_heading_bin=num_heading_bin, num_class=num_class, seed_feat_dim=256,
                                         use_kl_loss=self.use_kl_loss)
        if self.use_vote_weight:
            self.votes_weight_predictor = nn.Sequential(nn.Conv1d(256, 128, kernel_size=1), nn.BatchNorm1d(128),
                                                        nn.PReLU(), nn.Conv1d(128, 1, kernel_size=1), nn.Sigmoid())

    def forward(self, xyz, features, data_dict):
        """xyz (B, K, 3) votes, features (B, C, K) -> data_dict with the aggregated votes, the ROI head's
        scores and the decoded boxes (pred_center / pred_size / pred_heading / pred_bbox_corner)."""
        if self.use_vote_weight:
            vote_weights = self.votes_weight_predictor(features)
            data_dict["vote_weights"] = vote_weights
            features = features * vote_weights
        xyz, features, fps_inds = self.vote_aggregation(xyz, features)
        data_dict["aggregated_vote_xyz"] = xyz
        data_dict["aggregated_vote_features"] = features.permute(0, 2, 1).contiguous()
        data_dict["aggregated_vote_inds"] = fps_inds
        data_dict = self.proposal(features, data_dict)
        return self.decode_scores(data_dict)

    def decode_pred_box(self, data_dict):
        """:94-144, on the device."""
        data_dict = decode_pred_box(data_dict, self.num_heading_bin)
        if self.mask_box and self.training:  # :124-125; the corners follow the masked boxes
            from .box_util import get_3d_box_batch
            c, s = self.mask(data_dict["pred_center"], data_dict["pred_size"])
            data_dict["pred_center"], data_dict["pred_size"] = c, s
            data_dict["pred_bbox_corner"] = get_3d_box_batch(s.detach(), data_dict["pred_heading"].detach(), c.detach())
        return data_dict

    def decode_scores(self, data_dict):
        """:146-159."""
        data_dict = self.decode_pred_box(data_dict)
        data_dict["pred_bbox_feature"] = data_dict["aggregated_vote_features"]
        data_dict["pred_bbox_mask"] = data_dict["objectness_scores"].argmax(-1)
        data_dict["pred_bbox_sems"] = data_dict["sem_cls_scores"].argmax(-1)
        return data_dict

    def mask(self, pred_center, pred_box_size):
        """:161-178: replace 30 % of the boxes by random ones (training-time augmentation)."""
        batch_size, num_proposal, _ = pred_center.shape
        dev = pred_center.device
        mask_index = torch.bernoulli(torch.full([batch_size, num_proposal], 0.3)).bool().to(dev)
        m = mask_index[:, :, None].repeat(1, 1, 3)
        random_center = (torch.randn([batch_size, num_proposal, 3]) / 2).to(dev)
        random_size = (1 + torch.randn([batch_size, num_proposal, 3])).to(dev)
        return (~m) * pred_center + m * random_center, (~m) * pred_box_size + m * random_size


class VoteNetDetector(nn.Module):
    """The detection branch of JointNet.forward (models/jointnet/jointnet.py:138-154): backbone -> votes
    -> L2-normalised vote features -> ProposalModule, with the reference's attribute names
    (``backbone_net``, ``vgen``, ``proposal``) so those checkpoint entries load unchanged."""

    def __init__(self, num_class=18, num_heading_bin=1, num_size_cluster=18, mean_size_arr=None,
                 input_feature_dim=0, num_proposal=256, vote_factor=1, sampling="vote_fps", use_kl_loss=False):
        super().__init__()
        self.backbone_net = Pointnet2Backbone(input_feature_dim=input_feature_dim)
        self.vgen = VotingModule(vote_factor, 256)
        self.proposal = ProposalModule(num_class, num_heading_bin, num_size_cluster, mean_size_arr, num_proposal,
                                       sampling, use_kl_loss=use_kl_loss)

    def forward(self, data_dict):
        data_dict = self.backbone_net(data_dict)
        xyz, features = data_dict["fp2_xyz"], data_dict["fp2_features"]
        data_dict["seed_inds"] = data_dict["fp2_inds"]
        data_dict["seed_xyz"] = xyz
        data_dict["seed_features"] = features
        xyz, features = self.vgen(xyz, features)
        features_norm = torch.norm(features, p=2, dim=1)          # jointnet.py:148-149
        features = features.div(features_norm.unsqueeze(1))
        data_dict["vote_xyz"] = xyz
        data_dict["vote_features"] = features
        return self.proposal(xyz, features, data_dict)
