"""DIoU ("OID") reference loss, batched and sync-free.

Mirror of /root/reference/lib/loss_helper/loss_grounding.py:129-365 ``compute_diou_loss`` and
lib/loss_helper/loss.py:6-17 ``SoftmaxRankingLoss``.  The reference walks ``for i in range(batch_size):
for j in range(len_nun_max):`` in Python, calling box3d_diou_batch_tensor (~45 torch kernels) per
(scene, sentence) and reading ``max_ious`` / ``cnt`` back to the host for its ``if``s.  Here the IoU /
DIoU of all (scene, sentence, proposal) triples is one kernel (box_util.box3d_diou_pairs) and the label
logic is a handful of masked tensor ops: same labels, same losses, no host synchronisation
(``max_iou_rate_*`` stay 0-dim device tensors instead of numpy scalars).  Pinned against the reference's
own function, imported with stand-ins for the packages this image lacks, by
tests/golden/reference_losses.npz (tests/golden/make_golden_losses.py): eight configurations, values
and gradients.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .box_util import box3d_diou_batch_tensor, box3d_diou_pairs


class SoftmaxRankingLoss(nn.Module):
    """lib/loss_helper/loss.py:6-17."""

    def forward(self, inputs, targets):
        assert inputs.shape == targets.shape
        probs = F.softmax(inputs + 1e-8, dim=-1)
        return -torch.sum(torch.log(probs + 1e-8) * targets, dim=-1).mean()


def gt_boxes_from_labels(data_dict, config):
    """``config.param2obb_batch_tensor`` (data/scannet/model_util_scannet.py:177-190) for every scene at
    once: centre = the label, size = mean size of the class + residual (ScanNet headings are 0)."""
    center = data_dict['ref_center_label_list'].detach()[..., 0:3]
    size_class = data_dict['ref_size_class_label_list'].detach().long()
    size_residual = data_dict['ref_size_residual_label_list'].detach()
    mean = getattr(config, "mean_size_arr_tensor", None)
    if mean is None:
        mean = torch.as_tensor(config.mean_size_arr, dtype=torch.float32)
    mean = mean.to(device=center.device, dtype=center.dtype)
    return center, mean[size_class] + size_residual


def compute_diou_loss(data_dict, config, no_reference=False, use_reg_head=False, use_kl_loss=False, debug=False):
    """loss_grounding.py:129-365.  Returns (data_dict, ref_loss, cluster_preds (B,L,K), cluster_labels
    (B,L,K) float one-hot of the best-IoU proposal where it reaches 0.25); writes ``diou_loss``,
    ``max_iou_rate_0.25`` / ``_0.5`` (``kl_loss`` with use_kl_loss) and the statistics keys the reference
    solver reads unconditionally (lib/joint/solver_3dvlp.py:651-661,687-697: ``top_ind``,
    ``pred_iou_rate_*``, ``class_iou_rate_*``, ``class_size_*``, ``top_iou_rate_*`` -- zeros unless
    ``debug``) into data_dict like the reference.  Statistics are 0-dim device tensors where the reference
    stores numpy scalars: nothing here synchronises with the host."""
    pred_center = data_dict['pred_center']
    pred_size = data_dict['pred_size']
    dev = pred_center.device
    gt_center, gt_size = gt_boxes_from_labels(data_dict, config)          # (B,L,3)
    B, K = data_dict['aggregated_vote_features'].shape[:2]
    Lmax = gt_center.shape[1]
    lang_num = data_dict["lang_num"].to(dev)
    valid = torch.arange(Lmax, device=dev)[None, :] < lang_num[:, None]   # (B,L)
    cluster_preds = data_dict["cluster_ref"].reshape(B, Lmax, K)
    objectness_masks = data_dict['objectness_scores'].max(2)[1].float()    # (B,K)

    if use_reg_head:   # per-sentence refined boxes: (B,L,K,3) offsets on centre and size
        c1 = (pred_center[:, None] + data_dict['pred_center_reg']).reshape(-1, 3)
        s1 = (pred_size[:, None] + data_dict['pred_size_reg']).reshape(-1, 3)
        c2 = gt_center[:, :, None, :].expand(B, Lmax, K, 3).reshape(-1, 3)
        s2 = gt_size[:, :, None, :].expand(B, Lmax, K, 3).reshape(-1, 3)
        ious, dious = box3d_diou_batch_tensor(c1, s1, c2, s2)
        ious, dious = ious.view(B, Lmax, K), dious.view(B, Lmax, K)
    else:
        ious, dious = box3d_diou_pairs(pred_center, pred_size, gt_center, gt_size)

    ious_np = ious.detach()
    ious = ious_np  # only used for labels below (the reference's smooth_labels are detached too)
    if not no_reference and data_dict["random"] < 0.5:
        # loss_grounding.py:237-238: in training, half of the steps only rank proposals predicted as objects.
        # `istrain` may be a device tensor: folded into a select instead of read back
        istrain = data_dict["istrain"]
        istrain = istrain[0] if torch.is_tensor(istrain) or isinstance(istrain, (list, tuple)) else istrain
        if torch.is_tensor(istrain):
            ious = torch.where(istrain.to(dev) == 1, ious * objectness_masks[:, None, :], ious)
        elif istrain == 1:  # a host value: decided on the host, nothing uploaded (capturable in a CUDA graph)
            ious = ious * objectness_masks[:, None, :]

    ind_np = ious_np.argmax(-1)                                            # (B,L)
    max_ious = ious_np.gather(-1, ind_np.unsqueeze(-1)).squeeze(-1)
    hit25 = (max_ious >= 0.25) & valid
    hit5 = (max_ious >= 0.5) & valid
    labels = F.one_hot(ind_np, K).float() * hit25[..., None]

    ind_m = ious.argmax(-1)
    top = F.one_hot(ind_m, K).bool()
    if data_dict["epoch"] < 50:
        near = ious >= 0.25
        cnt = near.sum(-1, keepdim=True)
        multi = cnt >= 2
        smooth = torch.where(near & multi, 0.05 / (cnt - 1).clamp(min=1).float(), torch.zeros_like(ious))
        smooth = torch.where(top, torch.where(multi, 0.95, 1.0).to(ious.dtype).expand_as(ious), smooth)
    else:
        smooth = top.float()
    smooth = smooth * hit25[..., None]

    # SoftmaxRankingLoss per scene over its lang_num[i] sentences, then the batch mean
    logp = torch.log(F.softmax(cluster_preds + 1e-8, dim=-1) + 1e-8)
    per_row = -(logp * smooth).sum(-1) * valid
    loss = (per_row.sum(1) / lang_num.to(per_row.dtype)).sum() / B
    diou_loss = ((1 - dious) * smooth * valid[..., None]).sum() / B
    data_dict['diou_loss'] = diou_loss

    if use_kl_loss:  # loss_grounding.py:303-316 (including its size term, which compares pred with pred: zero)
        alpha = data_dict["alpha"]
        alpha_center, alpha_size = alpha[:, :, 0:3], alpha[:, :, 4:6]
        picked = torch.gather(pred_center.detach(), 1, ind_np.unsqueeze(-1).expand(-1, -1, 3))       # (B,L,3)
        sl1 = F.smooth_l1_loss(picked, gt_center, reduction='none') * valid[..., None]
        per_scene = sl1.sum((1, 2)) / (3 * lang_num.clamp(min=1)).to(sl1.dtype)                      # nn.SmoothL1Loss() mean
        kl = torch.sum(torch.exp(-alpha_center) * per_scene[:, None, None] + 0.5 * alpha_center) + torch.sum(0.5 * alpha_size)
        data_dict['kl_loss'] = kl / B

    total = lang_num.sum().to(torch.float32)
    data_dict['max_iou_rate_0.25'] = hit25.sum() / total
    data_dict['max_iou_rate_0.5'] = hit5.sum() / total
    if debug:
        _debug_statistics(data_dict, ious_np, cluster_preds, gt_size, valid, total)
    else:
        data_dict['top_ind'] = 0
        data_dict['pred_iou_rate_0.25'] = 0
        data_dict['pred_iou_rate_0.5'] = 0
        for i in range(0, 18):
            data_dict[f'class_iou_rate_{i}'] = 0
            data_dict[f'class_size_{i}'] = 0
        for i in range(1, 6):
            data_dict[f'top_iou_rate_{i}'] = 0
    return data_dict, loss, cluster_preds, labels.detach()


def _debug_statistics(data_dict, ious, cluster_preds, gt_size, valid, total):
    """The reference's ``debug=True`` bookkeeping (loss_grounding.py:170-180,263-278,293-299,325-340) for
    all (scene, sentence) pairs at once: rank of the predicted proposal in the IoU order, IoU hit rates,
    per-class IoU / box volume, the five best IoUs.  Ties in the IoU order (many proposals at 0) are
    resolved by index (a stable argsort -- what torch.sort on the CPU does; the reference does not say)."""
    B, Lmax, K = ious.shape
    vf = valid.to(ious.dtype)
    n_valid = vf.sum().clamp(min=1)
    sort_ious, sort_ind = torch.sort(ious, dim=-1, stable=True)                       # ascending
    pred = cluster_preds.argmax(-1)                                                   # (B,L)
    rank = (sort_ind == pred.unsqueeze(-1)).float().argmax(-1).to(ious.dtype)         # position of the prediction
    per_scene = (rank * vf).sum(1) / vf.sum(1).clamp(min=1)
    data_dict['top_ind'] = per_scene.mean() + 1
    data_dict['pred_iou_rate_0.25'] = (((ious >= 0.25).to(ious.dtype).mean(-1)) * vf).sum() / n_valid
    data_dict['pred_iou_rate_0.5'] = (((ious >= 0.5).to(ious.dtype).mean(-1)) * vf).sum() / n_valid
    top5 = (sort_ious[..., -5:] * vf[..., None]).sum((0, 1))                          # ascending: [-1] is the best
    for i in range(1, 6):
        data_dict[f'top_iou_rate_{i}'] = top5[5 - i] / total
    cat = data_dict["object_cat_list"].long().to(ious.device)[:, :Lmax].reshape(-1)
    flat_valid = vf.reshape(-1)
    iou_at_pred = ious.gather(-1, pred.unsqueeze(-1)).squeeze(-1).reshape(-1) * flat_valid
    volume = gt_size.prod(-1).reshape(-1) * flat_valid
    zeros = torch.zeros(18, device=ious.device, dtype=ious.dtype)
    class_iou = zeros.index_add(0, cat, iou_at_pred)
    class_size = zeros.index_add(0, cat, volume)
    class_cnt = zeros.index_add(0, cat, flat_valid)
    class_cnt = torch.where(class_cnt == 0, torch.ones_like(class_cnt), class_cnt)
    class_iou, class_size = class_iou / class_cnt, class_size / class_cnt
    for i in range(0, 18):
        data_dict[f'class_iou_rate_{i}'] = class_iou[i]
        data_dict[f'class_size_{i}'] = class_size[i]
