"""OCC / OSC contrastive module, batched, with optional cross-GPU negatives.

Mirror of /root/reference/models/constrast_module/constrast_module.py:9-131 (``create_box_batch``,
``SoftCrossEntropy``, ``NCELoss``, ``ContrastModule``; same parameter names, so reference checkpoints
load).  Differences in HOW, not WHAT:

* The reference gets the proposal / target IoU from ``pytorch3d.ops.box3d_overlap`` on corner boxes it
  has just built from centre + size, i.e. boxes that are axis-aligned by construction; the same IoU is
  one launch of csrc/box_ops.cu for all (scene, sentence, proposal) triples.  pytorch3d is not vendored
  or pinned by the reference, so this equivalence is pinned by the closed form, not by running it
  ("parity unpinned" for that third-party call).
* Its ``for i in range(batch_size): for j in range(len_nun_max):`` loops with ``torch.where`` on the
  objectness mask (a host sync per scene) become masked, batched softmaxes.
* ``forward(..., gather=True)`` is the north star's one collective: object and sentence embeddings
  of all ranks are all-gathered (``dist.gather_contrastive_embeddings``: one NCCL call, local slice
  keeps its autograd history) and every sentence is additionally contrasted against the objects of
  every other scene, which are negatives by construction.  This extends the reference, which only
  contrasts within a scene (constrast_module.py:85-126); ``gather=False`` is the reference semantics.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .box_util import box3d_diou_pairs
from .loss_grounding import gt_boxes_from_labels

_UNIT = ((-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1))


def create_box_batch(center, size):
    """constrast_module.py:9-16: (n,3) centre + size -> (n,8,3) corners (device follows the input)."""
    unit = torch.tensor(_UNIT, device=center.device, dtype=size.dtype)
    return (center[:, None, :] + size[:, None, :] * unit / 2).float()


def SoftCrossEntropy(inputs, target):
    """constrast_module.py:19-22."""
    return torch.mean(torch.mul(-F.log_softmax(inputs, dim=-1), target))


class NCELoss(nn.Module):
    """constrast_module.py:25-38 (the temperature is a parameter the reference never applies)."""

    def __init__(self, init_tau=0.07, clamp=4.6051):
        super().__init__()
        self.tau = nn.Parameter(torch.tensor([math.log(1.0 / init_tau)], dtype=torch.float32))
        self.clamp = clamp

    def forward(self, logits, iou_matrix):
        return (SoftCrossEntropy(logits, iou_matrix) + SoftCrossEntropy(logits.t(), iou_matrix)) / 2


def _masked_log_softmax(logits, mask):
    """log_softmax over the last dim restricted to `mask` (other entries come out as 0)."""
    neg = torch.finfo(logits.dtype).min
    x = torch.where(mask, logits, torch.full_like(logits, neg))
    return torch.where(mask, F.log_softmax(x, dim=-1), torch.zeros_like(logits))


def occ_osc_losses(text, box, box_iou, target, obj, valid, gather=False):
    """The two contrastive terms from normalised embeddings and IoU targets (pure tensor ops, any
    device): text (B,L,h) sentence embeddings, box / box_iou (B,K,h) object embeddings of the two
    projections, target (B,L,K) in {0,1} (proposal overlaps the sentence's object, kept proposals
    only), obj (B,K) bool kept proposals, valid (B,L) bool real sentences.
    Returns (lang_con_loss, iou_con_loss) as constrast_module.py:85-130 accumulates them.

    gather=True: the objects of every scene of every rank join the softmax of the sentence term as
    negatives (one all-gather of the embeddings; the local slice keeps its autograd history)."""
    B, K = obj.shape
    Lmax = text.shape[1]
    cnt = obj.sum(-1).clamp(min=1).to(text.dtype)                             # (B,)
    # OSC (sentence <-> objects).  NCELoss on a (1,cnt) row: the column direction is a softmax over
    # one element (= 0), so the loss is half the row cross-entropy averaged over the cnt entries.
    if gather:
        from . import dist as D
        rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        all_box, _ = D.gather_contrastive_embeddings(box.reshape(B * K, -1), text.reshape(B * Lmax, -1))
        all_obj = D.concat_all_gather(obj.reshape(B * K).to(torch.uint8)).bool()      # (world*B*K,)
        scenes = all_obj.numel() // K
        sim = torch.einsum("blh,nh->bln", text, all_box)
        logp = _masked_log_softmax(sim, all_obj[None, None, :].expand_as(sim)).view(B, Lmax, scenes, K)
        own = logp[torch.arange(B), :, rank * B + torch.arange(B, device=logp.device), :]   # (B,L,K)
        lang_rows = -(own * target).sum(-1) / cnt[:, None] / 2
    else:
        sim = torch.einsum("blh,bkh->blk", text, box)
        logp = _masked_log_softmax(sim, obj[:, None, :].expand_as(sim))
        lang_rows = -(logp * target).sum(-1) / cnt[:, None] / 2
    lang_con_loss = (lang_rows * valid).sum() / B

    # OCC (objects <-> objects): target = outer product of the IoU mask; both softmax directions of a
    # symmetric similarity matrix coincide -> one masked row softmax, mean over the cnt^2 entries
    sim_iou = torch.einsum("bkh,bmh->bkm", box_iou, box_iou)
    logp_iou = _masked_log_softmax(sim_iou, obj[:, :, None] & obj[:, None, :])
    tv = target * valid[..., None].to(target.dtype)
    quad = torch.einsum("blk,bkm,blm->b", tv, -logp_iou, tv)                  # sum over sentences, k, m
    iou_con_loss = (quad / (cnt * cnt)).sum() / B
    return lang_con_loss, iou_con_loss


class ContrastModule(nn.Module):
    def __init__(self, config, hidden=128):
        super().__init__()
        self.pc_proj = nn.Linear(hidden, hidden, bias=False)
        self.text_proj = nn.Linear(hidden, hidden, bias=False)
        self.nce_loss = NCELoss()
        self.config = config
        self.bce_loss = nn.BCEWithLogitsLoss()
        self.pc_proj_iou = nn.Sequential(nn.Linear(hidden, hidden, bias=False))

    def forward(self, data_dict, gather=False):
        if data_dict["epoch"] < 50:  # constrast_module.py:54-56
            data_dict["con_loss"] = torch.zeros(1)
            return data_dict
        pred_center = data_dict['pred_center'].detach()
        pred_size = data_dict['pred_size'].detach()
        features = data_dict["bbox_feature"]                                  # (B,K,hidden)
        B, K = data_dict['aggregated_vote_features'].shape[:2]
        gt_center, gt_size = gt_boxes_from_labels(data_dict, self.config)     # (B,L,3)
        Lmax = gt_center.shape[1]
        lang_num = data_dict["lang_num"].to(features.device)
        lang_emb = data_dict["lang_emb"].view(B, -1, data_dict["lang_emb"].shape[-1])[:, :Lmax]
        obj = data_dict['objectness_scores'].max(2)[1].bool()                 # (B,K) proposals kept
        valid = torch.arange(Lmax, device=features.device)[None, :] < lang_num[:, None]

        # IoU of every proposal with every target (target grown by 1e-2, constrast_module.py:100-101):
        # one launch instead of B x L calls of pytorch3d.box3d_overlap on axis-aligned corner boxes
        ious, _ = box3d_diou_pairs(pred_center, pred_size, gt_center, gt_size + 1e-2)
        target = ((ious > 0.25) & obj[:, None, :]).to(features.dtype).detach()

        text = F.normalize(self.text_proj(lang_emb), dim=-1)
        box = F.normalize(self.pc_proj(features), dim=-1)
        box_iou = F.normalize(self.pc_proj_iou(features), dim=-1)
        lang_con_loss, iou_con_loss = occ_osc_losses(text, box, box_iou, target, obj, valid, gather=gather)
        data_dict["lang_con_loss"] = lang_con_loss
        data_dict["iou_con_loss"] = iou_con_loss
        return data_dict
