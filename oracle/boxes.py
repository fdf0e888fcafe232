"""CPU oracle for the box / grounding-loss / contrastive rows -- TEST
INFRASTRUCTURE ONLY (see oracle/oracle.py for who may import this).

* ``box3d_diou_batch_tensor`` restates /root/reference/utils/box_util.py:488-528 op for op; it is pinned
  against the reference function itself by tests/golden/reference_boxes.npz
  (tests/golden/make_golden_boxes.py imports the reference in the build container).
* ``compute_diou_loss_loop`` restates lib/loss_helper/loss_grounding.py:129-365 with its Python loops; since
  round 2 it is pinned against the reference function itself (imported with stand-ins for torch_scatter /
  easydict by tests/golden/make_golden_losses.py): tests/test_losses_golden.py replays
  tests/golden/reference_losses.npz through it in the CPU suite.
* ``contrast_loop`` restates models/constrast_module/constrast_module.py:52-131 with its loops; pinned the same
  way against the reference's own ContrastModule.  pytorch3d's ``box3d_overlap`` (not vendored, not installed,
  version unpinned by the reference) is replaced -- here and in the generator's stand-in -- by the closed-form
  IoU of the axis-aligned boxes it is called on: that one third-party call stays "parity unpinned".
"""
import torch
import torch.nn.functional as F


def box3d_diou_batch_tensor(center1, size1, center2, size2):
    cx1, cy1, cz1 = center1[:, 0], center1[:, 1], center1[:, 2]
    cx2, cy2, cz2 = center2[:, 0], center2[:, 1], center2[:, 2]
    l1, w1, h1 = size1[:, 0], size1[:, 1], size1[:, 2]
    l2, w2, h2 = size2[:, 0], size2[:, 1], size2[:, 2]
    area1 = l1 * w1 * h1
    area2 = l2 * w2 * h2
    ixa, ixb = torch.max(cx1 - l1 / 2, cx2 - l2 / 2), torch.min(cx1 + l1 / 2, cx2 + l2 / 2)
    iya, iyb = torch.max(cy1 - w1 / 2, cy2 - w2 / 2), torch.min(cy1 + w1 / 2, cy2 + w2 / 2)
    iza, izb = torch.max(cz1 - h1 / 2, cz2 - h2 / 2), torch.min(cz1 + h1 / 2, cz2 + h2 / 2)
    inter = torch.clamp(ixb - ixa, min=0) * torch.clamp(iyb - iya, min=0) * torch.clamp(izb - iza, min=0)
    union = area1 + area2 - inter
    iou = inter / union
    inter_diag = (cx1 - cx2) ** 2 + (cy1 - cy2) ** 2 + (cz1 - cz2) ** 2
    oxa, oxb = torch.min(cx1 - l1 / 2, cx2 - l2 / 2), torch.max(cx1 + l1 / 2, cx2 + l2 / 2)
    oya, oyb = torch.min(cy1 - w1 / 2, cy2 - w2 / 2), torch.max(cy1 + w1 / 2, cy2 + w2 / 2)
    oza, ozb = torch.min(cz1 - h1 / 2, cz2 - h2 / 2), torch.max(cz1 + h1 / 2, cz2 + h2 / 2)
    outer_diag = torch.clamp(oxb - oxa, min=0) ** 2 + torch.clamp(oyb - oya, min=0) ** 2 + torch.clamp(ozb - oza, min=0) ** 2
    diou = iou - 1.5 * inter_diag / outer_diag
    return iou, torch.clamp(diou, min=-1, max=1)


def softmax_ranking_loss(inputs, targets):  # lib/loss_helper/loss.py:6-17
    probs = F.softmax(inputs + 1e-8, dim=-1)
    return -torch.sum(torch.log(probs + 1e-8) * targets, dim=-1).mean()


def compute_diou_loss_loop(d, mean_size_arr, no_reference=False, use_reg_head=False):
    """loss_grounding.py:129-365 (main path; debug and KL branches left out).  Returns
    (loss, diou_loss, gt_labels (B,L,K), rate25, rate5)."""
    pred_center, pred_size = d['pred_center'].clone(), d['pred_size'].clone()
    gt_center_list = d['ref_center_label_list']
    B, K = d['aggregated_vote_features'].shape[:2]
    Lmax = gt_center_list.shape[1]
    lang_num = d["lang_num"]
    cluster_preds = d["cluster_ref"].reshape(B, Lmax, K)
    r25 = r5 = 0
    loss = 0.
    diou_loss = 0.
    gt_labels = torch.zeros(B, Lmax, K)
    for i in range(B):
        objectness_masks = d['objectness_scores'].max(2)[1].float()
        gt_box_center = gt_center_list[i][:, 0:3]
        gt_box_size = mean_size_arr[d['ref_size_class_label_list'][i].long()] + d['ref_size_residual_label_list'][i]
        labels = torch.zeros(Lmax, K)
        smooth_labels = torch.zeros(Lmax, K)
        dious_labels = []
        for j in range(Lmax):
            if j < lang_num[i]:
                gs = gt_box_size[j][None, :].repeat(K, 1)
                gc = gt_box_center[j][None, :].repeat(K, 1)
                if use_reg_head:
                    ious, dious = box3d_diou_batch_tensor(pred_center[i] + d['pred_center_reg'][i][j],
                                                          pred_size[i] + d['pred_size_reg'][i][j], gc, gs)
                else:
                    ious, dious = box3d_diou_batch_tensor(pred_center[i], pred_size[i], gc, gs)
                ious_np = ious.detach()
                dious_labels.append(dious)
                if d["istrain"][0] == 1 and not no_reference and d["random"] < 0.5:
                    ious = ious * objectness_masks[i]
                ious_ind = ious_np.argmax()
                max_ious = ious_np[ious_ind]
                if max_ious >= 0.25:
                    labels[j, ious_ind] = 1
                    if d["epoch"] < 50:
                        smooth_mask = (ious >= 0.25)
                        cnt = torch.sum(smooth_mask)
                        if cnt >= 2:
                            smooth_labels[j, smooth_mask] = 0.05 / (cnt - 1)
                            smooth_labels[j, ious.argmax()] = 0.95
                        else:
                            smooth_labels[j, ious.argmax()] = 1
                    else:
                        smooth_labels[j, ious.argmax()] = 1
                    r25 += 1
                if max_ious >= 0.5:
                    r5 += 1
        cluster_labels = smooth_labels.detach()
        gt_labels[i] = labels
        n = int(lang_num[i])
        cluster_dious_labels = torch.stack(dious_labels).float()
        loss = loss + softmax_ranking_loss(cluster_preds[i, :n], cluster_labels[:n].float().clone())
        diou_loss = diou_loss + torch.sum((1 - cluster_dious_labels[:n]) * cluster_labels[:n].float().clone())
    total = float(torch.sum(lang_num))
    return loss / B, diou_loss / B, gt_labels, r25 / total, r5 / total


def _soft_ce(inputs, target):  # constrast_module.py:19-22
    return torch.mean(torch.mul(-F.log_softmax(inputs, dim=-1), target))


def _nce(logits, iou_matrix):  # constrast_module.py:32-38
    return (_soft_ce(logits, iou_matrix) + _soft_ce(logits.t(), iou_matrix)) / 2


def contrast_loop(d, mean_size_arr, w_pc, w_text, w_pc_iou):
    """constrast_module.py:52-131 with the closed-form axis-aligned IoU in place of pytorch3d.
    w_* are the three bias-free Linear weights.  Returns (lang_con_loss, iou_con_loss)."""
    pred_center, pred_size = d['pred_center'].detach(), d['pred_size'].detach()
    features = d["bbox_feature"]
    B, K = d['aggregated_vote_features'].shape[:2]
    Lmax = d['ref_center_label_list'].shape[1]
    lang_emb = d["lang_emb"].view(B, d["lang_emb"].shape[0] // B, -1)
    objectness_masks = d['objectness_scores'].max(2)[1].float()
    lang_con = 0.
    iou_con = 0.
    for i in range(B):
        gt_center = d['ref_center_label_list'][i][:, 0:3]
        gt_size = mean_size_arr[d['ref_size_class_label_list'][i].long()] + d['ref_size_residual_label_list'][i]
        object_index = torch.where(objectness_masks[i])[0]
        cnt = object_index.shape[0]
        fb = features[i][object_index]
        for j in range(Lmax):
            if j < d["lang_num"][i]:
                ious, _ = box3d_diou_batch_tensor(pred_center[i][object_index], pred_size[i][object_index],
                                                  gt_center[j][None].repeat(cnt, 1), (gt_size[j] + 1e-2)[None].repeat(cnt, 1))
                tm = (ious > 0.25).float().unsqueeze(0)
                text = F.normalize(lang_emb[i][j][None, :] @ w_text.t(), dim=-1)
                box = F.normalize(fb @ w_pc.t(), dim=-1)
                lang_con = lang_con + _nce(text @ box.t(), tm)
                tmi = (ious > 0.25).float()
                tmi = tmi.unsqueeze(0).repeat(cnt, 1) * tmi.unsqueeze(1).repeat(1, cnt)
                bi = F.normalize(fb @ w_pc_iou.t(), dim=-1)
                iou_con = iou_con + _nce(bi @ bi.t(), tmi)
    return lang_con / B, iou_con / B
