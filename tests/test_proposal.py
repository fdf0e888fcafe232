"""ProposalModule / StandardROIHeads / VoteNetDetector: CPU checks of the generic
path against the golden produced by the reference's own ProposalModule
(tests/golden/make_golden_proposal.py), GPU checks of the tensor-core fast path."""
import os

import numpy as np
import pytest
import torch

from make_golden_proposal import KEYS, fill_proposal, inputs

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "reference_proposal.npz"))


class _OracleExt:
    """Test-side stand-in for the CUDA extension: the oracle's nine reference ops plus the one fused
    entry point the module tree calls (FPS + xyz gather), composed from them."""

    def __init__(self, ext):
        self._ext = ext

    def __getattr__(self, name):
        return getattr(self._ext, name)

    def furthest_point_sampling_gather(self, points, nsamples, ordered=False):
        inds = self._ext.furthest_point_sampling(points, nsamples)
        new_xyz = torch.gather(points, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        return inds, new_xyz


def _module(pkg, nh, kl, num_proposal=64):
    import importlib
    prop = importlib.import_module("3dvlp_b200.proposal")
    pm = prop.ProposalModule(18, nh, 18, np.ones((18, 3), dtype=np.float32), num_proposal, "vote_fps", use_kl_loss=kl)
    fill_proposal(pm)
    return pm.eval()


@pytest.mark.parametrize("tag,nh,kl", [("nh1", 1, False), ("nh4_kl", 4, True)])
def test_proposal_generic_path_matches_reference(pkg, oracle, gold, monkeypatch, tag, nh, kl):
    """CPU: our module tree over the oracle ops == the reference's ProposalModule over the same ops."""
    shim = _OracleExt(oracle.ext)
    monkeypatch.setattr(pkg.pointnet2_utils, "_ext", shim)
    monkeypatch.setattr(pkg.pointnet2_modules.pointnet2_utils, "_ext", shim)
    pm = _module(pkg, nh, kl)
    assert sorted(pm.state_dict().keys()) == list(gold[f"{tag}_state_keys"])
    votes, feats = inputs()
    with torch.no_grad():
        d = pm(votes, feats, {})
    for k in KEYS + (("alpha",) if kl else ()):
        want = gold[f"{tag}_{k}"]
        got = d[k].numpy()
        assert got.shape == want.shape, k
        if want.dtype.kind in "iu":
            assert np.array_equal(got, want), k
        else:
            np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6, err_msg=k)


def test_roi_head_parameter_layout(pkg):
    import importlib
    prop = importlib.import_module("3dvlp_b200.proposal")
    det = prop.VoteNetDetector(input_feature_dim=1, mean_size_arr=np.ones((18, 3)))
    keys = det.state_dict().keys()
    for k in ("backbone_net.fp2.mlp.layer1.bn.bn.running_var", "vgen.conv3.bias",
              "proposal.vote_aggregation.mlp_module.layer0.conv.weight", "proposal.proposal.convs.4.running_mean",
              "proposal.proposal.box_predictor.weight", "proposal.proposal.heading_reg_predictor.bias"):
        assert k in keys, k
    head = prop.StandardROIHeads(1, 18)
    assert head.objectness_predictor.weight.std().item() < 5e-3      # roi_heads.py:72-75
    with pytest.raises(NotImplementedError):
        prop.StandardROIHeads(1, 18, back_trace_type="x")


@pytest.mark.gpu
@pytest.mark.parametrize("tag,nh,kl", [("nh1", 1, False), ("nh4_kl", 4, True)])
def test_proposal_fused_path(pkg, gold, cuda, tag, nh, kl):
    """GPU eval path (fused SA kernel -> three tcgen05 linears -> on-device box decode) against the
    reference's fp32 ProposalModule: indices bit-exact, floats within the bf16-operand contract."""
    pm = _module(pkg, nh, kl).to(cuda)
    votes, feats = inputs()
    with torch.no_grad():
        d = pm(votes.to(cuda), feats.to(cuda), {})
    assert pm.proposal._fused is not None, "tensor-core path was not taken"
    assert np.array_equal(d["aggregated_vote_inds"].cpu().numpy(), gold[f"{tag}_aggregated_vote_inds"])
    assert np.array_equal(d["aggregated_vote_xyz"].cpu().numpy(), gold[f"{tag}_aggregated_vote_xyz"])
    for k in ("aggregated_vote_features", "objectness_scores", "sem_cls_scores", "heading_scores",
              "heading_residuals", "rois", "pred_size", "pred_center") + (("alpha",) if kl else ()):
        want = gold[f"{tag}_{k}"]
        got = d[k].float().cpu().numpy()
        assert got.shape == want.shape, k
        scale = max(float(np.abs(want).max()), 1e-6)
        assert float(np.abs(got - want).max()) < 5e-2 * scale, (k, float(np.abs(got - want).max()), scale)
    # labels may flip only where the two logits are within the contract's error of each other
    for k, src in (("bbox_mask", "objectness_scores"), ("pred_bbox_sems", "sem_cls_scores")):
        want, got = gold[f"{tag}_{k}"], d[k].cpu().numpy()
        s = np.sort(gold[f"{tag}_{src}"], -1)
        margin = s[..., -1] - s[..., -2]
        bad = (got != want) & (margin > 1e-1 * np.abs(gold[f"{tag}_{src}"]).max())
        assert not bad.any(), k
    # corners are consistent with the decoded centre / size / heading (exact functional relation)
    from importlib import import_module
    bu = import_module("3dvlp_b200.box_util")
    c = bu.get_3d_box_batch(d["pred_size"], d["pred_heading"], d["pred_center"])
    assert torch.equal(c, d["pred_bbox_corner"])
    # and the generic path on the GPU (fp32, no TF32) reproduces the golden tightly
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.enable_grad():
            w = pm(votes.to(cuda), feats.to(cuda), {})
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    for k in ("objectness_scores", "rois", "pred_center", "pred_bbox_corner"):
        np.testing.assert_allclose(w[k].detach().cpu().numpy(), gold[f"{tag}_{k}"], rtol=2e-4, atol=2e-4, err_msg=k)


@pytest.mark.gpu
def test_votenet_detector_graph_capturable(pkg, cuda):
    """The whole detection branch (backbone -> votes -> proposals -> boxes) has no host sync: it can be
    captured into a CUDA graph and replays to the eager result."""
    import importlib
    from common import fill_state_dict, fill_state_dict_conv1d
    prop = importlib.import_module("3dvlp_b200.proposal")
    syn = importlib.import_module("3dvlp_b200.synthetic")
    det = prop.VoteNetDetector(input_feature_dim=1, mean_size_arr=np.ones((18, 3)))
    fill_state_dict(det.backbone_net)
    fill_state_dict_conv1d(det.vgen)
    fill_proposal(det.proposal)
    det = det.to(cuda).eval()
    pc = syn.make_batch(2, 2, 9000, 1).to(cuda)
    with torch.no_grad():
        want = det({"point_clouds": pc})
    g = pkg.GraphedBackbone(det, pc)
    assert g.graph is not None, g.capture_error
    out = g(pc)
    torch.cuda.synchronize()
    assert out["pred_bbox_corner"].shape == (2, 256, 8, 3)
    for k in ("aggregated_vote_inds", "objectness_scores", "rois", "pred_center", "pred_bbox_corner", "pred_bbox_sems"):
        assert torch.equal(out[k], want[k]), k
