"""Vote / objectness losses on the fused nn_distance kernel vs a literal
restatement of lib/loss_helper/loss_detection.py:24-113 built on the reference's own nn_distance
arithmetic (oracle.nn_distance_torch) on CPU -- values and gradients, at the live shapes
(B*1024, 1, 3) x (B*1024, 3, 3) l1 and (B, 256, 3) x (B, 256, 3) L2."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_vote_loss(O, d):
    b, n = d["seed_xyz"].shape[:2]
    inds = d["seed_inds"].long()
    mask = torch.gather(d["vote_label_mask"], 1, inds)
    gt = torch.gather(d["vote_label"], 1, inds.view(b, n, 1).repeat(1, 1, 9)) + d["seed_xyz"].repeat(1, 1, 3)
    _, _, dist2, _ = O.nn_distance_torch(d["vote_xyz"].view(b * n, -1, 3), gt.view(b * n, 3, 3), l1=True)
    vd = torch.min(dist2, dim=1)[0].view(b, n)
    return torch.sum(vd * mask.float()) / (torch.sum(mask.float()) + 1e-6)


def _ref_objectness(O, d):
    dist1, ind1, _, _ = O.nn_distance_torch(d["aggregated_vote_xyz"], d["center_label"][:, :, 0:3])
    e = torch.sqrt(dist1 + 1e-6)
    label = (e < 0.3).long()
    mask = ((e < 0.3) | (e > 0.3)).float()
    loss = torch.nn.functional.cross_entropy(d["objectness_scores"].transpose(2, 1), label,
                                             weight=torch.tensor([0.2, 0.8]), reduction="none")
    return torch.sum(loss * mask) / (torch.sum(mask) + 1e-6), label, mask, ind1


def _batch(g, b=4, npts=5000, nseed=1024, k=256):
    d = {
        "seed_xyz": torch.rand(b, nseed, 3, generator=g) * 4 - 2,
        "seed_inds": torch.randint(0, npts, (b, nseed), generator=g, dtype=torch.int32),
        "vote_label_mask": (torch.rand(b, npts, generator=g) < 0.4).long(),
        "vote_label": torch.randn(b, npts, 9, generator=g) * 0.3,
        "aggregated_vote_xyz": torch.rand(b, k, 3, generator=g) * 4 - 2,
        "center_label": torch.zeros(b, k, 3),
        "objectness_scores": torch.randn(b, k, 2, generator=g),
    }
    d["center_label"][:, :20] = torch.rand(b, 20, 3, generator=g) * 4 - 2  # 236 padded all-zero GT slots
    d["vote_xyz"] = d["seed_xyz"] + 0.2 * torch.randn(b, nseed, 3, generator=g)
    return d


def test_vote_and_objectness_losses(pkg, oracle, cuda):
    ld = __import__("importlib").import_module("3dvlp_b200.loss_detection")
    g = torch.Generator().manual_seed(11)
    d = _batch(g)
    d["vote_xyz"].requires_grad_(True)
    want_v = _ref_vote_loss(oracle, d)
    want_v.backward()
    want_o, wl, wm, wa = _ref_objectness(oracle, d)

    dg = {k: v.detach().to(cuda) for k, v in d.items()}
    dg["vote_xyz"].requires_grad_(True)
    got_v = ld.compute_vote_loss(dg)
    got_v.backward()
    torch.testing.assert_close(got_v.cpu(), want_v.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dg["vote_xyz"].grad.cpu(), d["vote_xyz"].grad, rtol=1e-4, atol=1e-7)
    got_o, gl, gm, ga = ld.compute_objectness_loss(dg)
    torch.testing.assert_close(got_o.cpu(), want_o, rtol=1e-5, atol=1e-6)
    assert torch.equal(gl.cpu(), wl) and torch.equal(gm.cpu(), wm) and torch.equal(ga.cpu(), wa)
    assert ga.dtype == torch.int64
