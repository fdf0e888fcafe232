"""Golden vectors for the box / loss building blocks of SURVEY.md 8f ranks 3-4, produced by the
REFERENCE's own functions (run once in the build container; needs /root/reference):

    python tests/golden/make_golden_boxes.py

* utils/box_util.py: box3d_diou_batch_tensor (values and torch-autograd gradients w.r.t. all four
  inputs), get_3d_box_batch (numpy), rotz_batch_pytorch
* lib/loss_helper/loss.py: SoftmaxRankingLoss
* decode_pred_box: ProposalModule itself cannot be imported here (easydict missing), so its body
  (proposal_module_fcos.py:94-139) is replayed on top of the reference's rotz_batch_pytorch /
  get_3d_box_batch.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from common import det_tensor  # noqa: E402


def box_cases():
    """(P,3) x 4 deterministic pairs covering overlap, containment, disjoint, touching faces,
    identical boxes (ties in every max/min) and shared faces."""
    P = 400
    c1 = det_tensor("c1", (P, 3), 1.5)
    s1 = det_tensor("s1", (P, 3), 0.9).abs() + 0.15
    c2 = c1 + det_tensor("dc", (P, 3), 0.8)
    s2 = det_tensor("s2", (P, 3), 0.9).abs() + 0.15
    c2[0:20] = c1[0:20]; s2[0:20] = s1[0:20]                      # identical
    c2[20:40] = c1[20:40] + 5.0                                   # disjoint
    s2[40:60] = s1[40:60] * 0.25; c2[40:60] = c1[40:60]           # contained, concentric
    c2[60:80, 0] = c1[60:80, 0] + (s1[60:80, 0] + s2[60:80, 0]) / 2   # touching along x
    c2[60:80, 1:] = c1[60:80, 1:]
    c2[80:100, 0] = c1[80:100, 0]; s2[80:100, 0] = s1[80:100, 0]  # same extent on one axis
    return c1.contiguous(), s1.contiguous(), c2.contiguous(), s2.contiguous()


def main():
    torch.set_num_threads(1)
    sys.path.insert(0, "/root/reference")  # only the generator reads the reference
    import utils.box_util as bu
    from lib.loss_helper.loss import SoftmaxRankingLoss
    out = {}
    c1, s1, c2, s2 = [t.clone().requires_grad_(True) for t in box_cases()]
    iou, diou = bu.box3d_diou_batch_tensor(c1, s1, c2, s2)
    gi, gd = det_tensor("gi", iou.shape, 1.0), det_tensor("gd", diou.shape, 1.0)
    (iou * gi).sum().backward(retain_graph=True)
    out["diou_iou"], out["diou_diou"] = iou.detach().numpy(), diou.detach().numpy()
    for n, t in zip(("c1", "s1", "c2", "s2"), (c1, s1, c2, s2)):
        out["diou_giou_" + n] = t.grad.numpy().copy()
        t.grad = None
    (diou * gd).sum().backward()
    for n, t in zip(("c1", "s1", "c2", "s2"), (c1, s1, c2, s2)):
        out["diou_gdiou_" + n] = t.grad.numpy().copy()

    size = det_tensor("bs", (2, 7, 3), 1.0).abs() + 0.2
    heading = det_tensor("bh", (2, 7), 3.0)
    center = det_tensor("bc", (2, 7, 3), 2.0)
    out["corners"] = bu.get_3d_box_batch(size.numpy(), heading.numpy(), center.numpy())
    out["rotz"] = bu.rotz_batch_pytorch(heading).numpy()

    x, t = det_tensor("rank_x", (5, 16), 3.0), det_tensor("rank_t", (5, 16), 0.5).abs()
    out["softmax_ranking"] = SoftmaxRankingLoss()(x, t).numpy()

    # decode_pred_box body (proposal_module_fcos.py:94-139), num_heading_bin = 1 as in ScanNet
    B, K, NH = 2, 9, 4
    xyz = det_tensor("dx", (B, K, 3), 2.0)
    hs, hr = det_tensor("dhs", (B, K, NH), 1.0), det_tensor("dhr", (B, K, NH), 0.3)
    rois = det_tensor("droi", (B, K, 6), 0.7).abs() + 0.05
    cls = torch.argmax(hs, -1)
    res = torch.gather(hr, 2, cls.unsqueeze(-1))
    pred_heading = cls.float() * (2.0 * np.pi / NH) + res[..., 0]
    pred_box_size = rois[:, :, 0:3] + rois[:, :, 3:6]
    vote_xyz = (rois[:, :, 0:3] - rois[:, :, 3:6]) / 2
    R = bu.rotz_batch_pytorch(pred_heading.float()).view(-1, 3, 3)
    vote_xyz = torch.matmul(vote_xyz.reshape(-1, 3).unsqueeze(1), R).squeeze(2).view(B, K, 3)
    pred_center = xyz - vote_xyz
    out["dec_heading"], out["dec_size"], out["dec_center"] = pred_heading.numpy(), pred_box_size.numpy(), pred_center.numpy()
    out["dec_corners"] = bu.get_3d_box_batch(pred_box_size.numpy(), pred_heading.numpy(), pred_center.numpy()).astype(np.float32)

    np.savez_compressed(os.path.join(HERE, "reference_boxes.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
