"""Deterministic inputs for the loss rows of SURVEY.md 8(f) (detection losses, DIoU reference loss,
OCC/OSC contrastive module), shared by make_golden_losses.py -- which feeds them to the REFERENCE's own
functions in the build container -- and by the tests that replay the fixture on the oracle / the GPU.
numpy's MT19937 only (stable across versions and machines)."""
import types

import numpy as np
import torch

from common import det_tensor

MEAN_SIZES = det_tensor("mean_sizes", (18, 3), 0.6).abs() + 0.3
NUM_CLASS = 18


def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def grounding_inputs(seed, B=3, K=32, Lmax=5, epoch=10, istrain=1, rnd=0.3, peaked_preds=False):
    """A data_dict shaped like the one compute_diou_loss / ContrastModule read
    (lib/loss_helper/loss_grounding.py:129-186, models/constrast_module/constrast_module.py:52-84)."""
    rs = np.random.RandomState(seed)
    gt_center = rs.uniform(0, 3, (B, Lmax, 3)).astype(np.float32)
    size_cls = rs.randint(0, 18, (B, Lmax))
    size_res = ((rs.uniform(size=(B, Lmax, 3)) - 0.5) * 0.1).astype(np.float32)
    gt_size = MEAN_SIZES.numpy()[size_cls] + size_res
    pred_center = rs.uniform(0, 3, (B, K, 3)).astype(np.float32)
    pred_size = (rs.uniform(size=(B, K, 3)) * 0.8 + 0.2).astype(np.float32)
    # plant proposals on the targets: several good ones for sentence 0, one for sentence 1, two for sentence 2 ...
    planted = ((0, 4), (1, 1), (2, 2), (3, 1), (4, 1))[:Lmax]
    for b in range(B):
        for j, n_good in planted:
            for t in range(n_good):
                k = 5 * j + t
                pred_center[b, k] = gt_center[b, j] + (rs.uniform(size=3) - 0.5) * 0.08
                pred_size[b, k] = gt_size[b, j] * (1 + (rs.uniform(size=3) - 0.5) * 0.1)
    obj = rs.normal(size=(B, K, 2)).astype(np.float32)
    obj[:, :12, 1] += 3.0  # most planted proposals count as objects
    cluster_ref = rs.normal(size=(B * Lmax, K)).astype(np.float32)
    if peaked_preds:
        # the matching head points at a planted proposal of its sentence, whose IoU is unique in its row: the
        # reference's debug statistic `top_ind` (rank of the prediction in an argsort of the IoUs) is only
        # defined up to the sort's tie order otherwise (most proposals tie at IoU 0)
        for b in range(B):
            for j, _ in planted:
                cluster_ref[b * Lmax + j, 5 * j] += 8.0
    d = {
        "pred_center": _t(pred_center), "pred_size": _t(pred_size),
        "pred_heading": torch.zeros(B, K),
        "pred_center_reg": _t((rs.uniform(size=(B, Lmax, K, 3)) - 0.5) * 0.04),
        "pred_size_reg": _t((rs.uniform(size=(B, Lmax, K, 3)) - 0.5) * 0.04),
        "ref_center_label_list": _t(gt_center), "ref_size_class_label_list": _t(size_cls, torch.long),
        "ref_size_residual_label_list": _t(size_res),
        "ref_heading_class_label_list": torch.zeros(B, Lmax, dtype=torch.long),
        "ref_heading_residual_label_list": torch.zeros(B, Lmax),
        "aggregated_vote_features": torch.zeros(B, K, 128),
        "lang_num": torch.tensor([Lmax, 3, 1, 2][:B]),
        "cluster_ref": _t(cluster_ref),
        "objectness_scores": _t(obj),
        "alpha": _t(rs.uniform(-0.05, 0.05, (B, K, 6))),
        "object_cat_list": _t(rs.randint(0, 18, (B, Lmax)), torch.long),
        "istrain": torch.tensor([istrain]), "random": rnd, "epoch": epoch,
        "bbox_feature": _t(rs.normal(size=(B, K, 128))),
        "lang_emb": _t(rs.normal(size=(B * Lmax, 128))),
    }
    return d


GROUNDING_CASES = {
    # tag: (seed, kwargs of grounding_inputs, kwargs of compute_diou_loss)
    "early_masked": (31, dict(epoch=10, istrain=1, rnd=0.3), {}),
    "early_plain": (32, dict(epoch=10, istrain=1, rnd=0.9), {}),
    "late_eval": (33, dict(epoch=60, istrain=0, rnd=0.3), {}),
    "early_reg": (34, dict(epoch=10, istrain=1, rnd=0.3), dict(use_reg_head=True)),
    "late_reg": (35, dict(epoch=60, istrain=1, rnd=0.2), dict(use_reg_head=True)),
    "early_kl": (36, dict(epoch=10, istrain=1, rnd=0.7), dict(use_kl_loss=True)),
    "no_reference": (37, dict(epoch=10, istrain=1, rnd=0.1), dict(no_reference=True)),
    "debug": (38, dict(epoch=10, istrain=0, rnd=0.6, peaked_preds=True), dict(debug=True)),
}


def detection_inputs(seed, B=3, npts=2000, nseed=128, K=64, K2=16, n_gt=6):
    """A data_dict shaped like the one lib/loss_helper/loss_detection.py reads (vote / objectness / box /
    semantic-class losses): K proposals, K2 padded ground-truth slots of which n_gt are real."""
    rs = np.random.RandomState(seed)
    seed_xyz = rs.uniform(-2, 2, (B, nseed, 3)).astype(np.float32)
    center = np.zeros((B, K2, 3), np.float32)
    center[:, :n_gt] = rs.uniform(-2, 2, (B, n_gt, 3))
    agg = rs.uniform(-2, 2, (B, K, 3)).astype(np.float32)
    for b in range(B):  # half of the proposals sit near a real object (inside and outside its box)
        for k in range(K // 2):
            agg[b, k] = center[b, k % n_gt] + rs.normal(scale=0.12, size=3)
    size_cls = rs.randint(0, 18, (B, K2))
    d = {
        "seed_xyz": _t(seed_xyz),
        "seed_inds": _t(rs.randint(0, npts, (B, nseed)), torch.int32),
        "vote_label_mask": _t(rs.uniform(size=(B, npts)) < 0.4, torch.long),
        "vote_label": _t(rs.normal(scale=0.3, size=(B, npts, 9))),
        "vote_xyz": _t(seed_xyz + rs.normal(scale=0.2, size=(B, nseed, 3))),
        "aggregated_vote_xyz": _t(agg),
        "center_label": _t(center),
        "heading_class_label": torch.zeros(B, K2, dtype=torch.long),
        "heading_residual_label": _t(rs.normal(scale=0.05, size=(B, K2))),
        "size_class_label": _t(size_cls, torch.long),
        "size_residual_label": _t((rs.uniform(size=(B, K2, 3)) - 0.5) * 0.1),
        "sem_cls_label": _t(rs.randint(0, NUM_CLASS, (B, K2)), torch.long),
        "objectness_scores": _t(rs.normal(size=(B, K, 2))),
        "heading_scores": _t(rs.normal(size=(B, K, 1))),
        "heading_residuals_normalized": _t(rs.normal(scale=0.3, size=(B, K, 1))),
        "rois": _t(np.abs(rs.normal(scale=0.4, size=(B, K, 6))) + 0.05),
        "sem_cls_scores": _t(rs.normal(size=(B, K, NUM_CLASS))),
    }
    return d


def make_config(cls=None):
    """The dataset config the losses read.  `cls` = the reference's ScannetDatasetConfig class (its
    __init__ reads dataset files that do not ship, so the instance is made without it and given the
    attributes the loss code touches); None -> a plain namespace with the same attributes."""
    cfg = types.SimpleNamespace() if cls is None else cls.__new__(cls)
    cfg.num_class = NUM_CLASS
    cfg.num_heading_bin = 1
    cfg.num_size_cluster = NUM_CLASS
    cfg.mean_size_arr = MEAN_SIZES.numpy().copy()
    cfg.mean_size_arr_tensor = MEAN_SIZES.clone()
    return cfg


def clone_dict(d, grad=()):
    out = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()}
    for k in grad:
        out[k].requires_grad_(True)
    return out
