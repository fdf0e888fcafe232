"""Golden vectors for the proposal stage, produced by the REFERENCE's own
``ProposalModule`` / ``StandardROIHeads`` (models/proposal_module/proposal_module_fcos.py,
models/proposal_module/ROI_heads/roi_heads.py), imported unmodified from /root/reference:

    python tests/golden/make_golden_proposal.py          (build container only)

Two things the import needs that this image lacks: ``easydict`` (pulled in by lib/config.py through
data/scannet/model_util_scannet.py) -- a ten-line attribute-dict stand-in is registered for the import --
and the compiled ``pointnet2._ext``, replaced by the C oracle as in make_golden.py.  Everything that
computes the fixture is the reference's code: vote aggregation (PointnetSAModuleVotes), the ROI head,
decode_pred_box (including its ``.cpu().numpy()`` corner decode) and decode_scores.
"""
import builtins
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from common import cloud, fill_state_dict, fill_state_dict_conv1d  # noqa: E402
from oracle import oracle as O  # noqa: E402

KEYS = ("aggregated_vote_xyz", "aggregated_vote_features", "aggregated_vote_inds", "objectness_scores", "rois",
        "heading_scores", "heading_residuals_normalized", "heading_residuals", "sem_cls_scores", "bbox_mask",
        "pred_heading", "pred_size", "pred_center", "pred_bbox_corner", "pred_bbox_mask", "pred_bbox_sems")


def fill_proposal(pm):
    """Deterministic parameters for a ProposalModule (reference or ours: same state-dict keys)."""
    fill_state_dict(pm.vote_aggregation)
    fill_state_dict_conv1d(pm.proposal)
    sd = pm.proposal.state_dict()  # predictors are initialised with std 1e-3 in the reference; keep the
    for k in sd:                   # golden sensitive by using O(1) outputs, but a tame exp() for the rois
        if k.startswith("box_predictor.weight"):
            sd[k] = sd[k] * 0.3
    pm.proposal.load_state_dict(sd)


def inputs():
    votes = cloud(71, 2, 300, c_extra=0, extent=1.0, dup_frac=0.05)[..., :3].contiguous()
    feats = cloud(72, 2, 300, c_extra=253)[..., :256].transpose(1, 2).contiguous() - 0.4
    feats = feats / feats.norm(p=2, dim=1, keepdim=True)   # votes arrive L2-normalised (jointnet.py:148-149)
    return votes, feats.contiguous()


def main():
    torch.set_num_threads(1)

    class EasyDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    stub = types.ModuleType("easydict")
    stub.EasyDict = EasyDict
    sys.modules["easydict"] = stub
    builtins.__POINTNET2_SETUP__ = True
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "lib", "pointnet2"))
    import pointnet2_utils as ref_utils
    ref_utils._ext = O.ext
    import lib.pointnet2.pointnet2_utils as ref_utils2
    ref_utils2._ext = O.ext
    from models.proposal_module.proposal_module_fcos import ProposalModule as RefProposal

    out = {}
    votes, feats = inputs()
    for tag, nh, kl in (("nh1", 1, False), ("nh4_kl", 4, True)):
        pm = RefProposal(18, nh, 18, np.ones((18, 3), dtype=np.float32), 64, "vote_fps", use_kl_loss=kl)
        fill_proposal(pm)
        pm.eval()
        with torch.no_grad():
            d = pm(votes, feats, {})
        for k in KEYS + (("alpha",) if kl else ()):
            v = d[k]
            out[f"{tag}_{k}"] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
        out[f"{tag}_state_keys"] = np.array(sorted(pm.state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, "reference_proposal.npz"), **out)
    print("wrote", len(out), "arrays", {k: v.shape for k, v in out.items() if k.startswith("nh1_")})


if __name__ == "__main__":
    main()
