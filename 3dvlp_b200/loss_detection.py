"""Vote and objectness losses -- the two live call sites of nn_distance.

Mirror of /root/reference/lib/loss_helper/loss_detection.py:24-113 (``compute_vote_loss``,
``compute_objectness_loss``).  Same data_dict keys, constants and
return values; the pairwise matching runs on the fused nn_distance kernel (one launch per direction,
differentiable) instead of the reference's repeat/sub/abs/sum/min chain, and nothing is pinned to
``.cuda()`` (the reference hard-codes it): tensors follow the inputs' device.
"""
import torch
import torch.nn.functional as F

from .nn_distance import nn_distance

FAR_THRESHOLD = 0.3
NEAR_THRESHOLD = 0.3
GT_VOTE_FACTOR = 3  # number of GT votes per point
OBJECTNESS_CLS_WEIGHTS = [0.2, 0.8]  # put larger weights on positive objectness


def compute_vote_loss(data_dict):
    """Match every seed's predicted votes to its (up to 3) ground-truth votes: L1 distance, min over
    the vote x GT-vote pairs, averaged over seeds that lie on an object (loss_detection.py:24-71)."""
    seed_xyz = data_dict['seed_xyz']
    batch_size, num_seed = seed_xyz.shape[0], seed_xyz.shape[1]
    vote_xyz = data_dict['vote_xyz']                      # (B, num_seed * vote_factor, 3)
    seed_inds = data_dict['seed_inds'].long()             # (B, num_seed) into the input cloud

    seed_gt_votes_mask = torch.gather(data_dict['vote_label_mask'], 1, seed_inds)
    gather_idx = seed_inds.view(batch_size, num_seed, 1).expand(-1, -1, 3 * GT_VOTE_FACTOR)
    seed_gt_votes = torch.gather(data_dict['vote_label'], 1, gather_idx) + seed_xyz.repeat(1, 1, GT_VOTE_FACTOR)

    votes = vote_xyz.reshape(batch_size * num_seed, -1, 3)
    gt_votes = seed_gt_votes.reshape(batch_size * num_seed, GT_VOTE_FACTOR, 3)
    # a vote to nowhere is not penalised as long as one vote is near a GT vote: min over dist2
    _, _, dist2, _ = nn_distance(votes, gt_votes, l1=True)
    votes_dist = torch.min(dist2, dim=1)[0].view(batch_size, num_seed)
    mask = seed_gt_votes_mask.float()
    return torch.sum(votes_dist * mask) / (torch.sum(mask) + 1e-6)


def compute_objectness_loss(data_dict):
    """Objectness labels from the distance of every proposal to its nearest GT centre, weighted
    cross-entropy over the non-gray-zone proposals (loss_detection.py:73-113).
    Returns (objectness_loss, objectness_label (B,K) long, objectness_mask (B,K) float,
    object_assignment (B,K) long in [0, K2))."""
    aggregated_vote_xyz = data_dict['aggregated_vote_xyz']
    gt_center = data_dict['center_label'][:, :, 0:3]
    dist1, ind1, _, _ = nn_distance(aggregated_vote_xyz, gt_center)   # (B,K): squared L2 to nearest GT

    euclidean_dist1 = torch.sqrt(dist1 + 1e-6)
    near = euclidean_dist1 < NEAR_THRESHOLD
    objectness_label = near.long()
    objectness_mask = (near | (euclidean_dist1 > FAR_THRESHOLD)).float()

    objectness_scores = data_dict['objectness_scores']
    weight = torch.tensor(OBJECTNESS_CLS_WEIGHTS, device=objectness_scores.device, dtype=objectness_scores.dtype)
    loss = F.cross_entropy(objectness_scores.transpose(2, 1), objectness_label, weight=weight, reduction='none')
    objectness_loss = torch.sum(loss * objectness_mask) / (torch.sum(objectness_mask) + 1e-6)
    return objectness_loss, objectness_label, objectness_mask, ind1
