"""BASELINE.json configs 4 and 5 at full size (40k points x 132 input channels; 20k/50k/100k-point
scenes, SA1 npoint 4096, batch 32): the oracle where it finishes in seconds, otherwise size-independent
properties -- the fused eval path against the generic op-by-op path on the same device, batch
invariance, index range / uniqueness."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(seed, b, n, c_in):
    syn = importlib.import_module("3dvlp_b200.synthetic")
    return syn.make_batch(seed, b, n, c_in)


def test_fps_and_ball_query_100k_points_vs_oracle(pkg, oracle, cuda):
    x = _scene(3, 1, 100000, 0)[..., :3].contiguous()
    want = oracle.ext.furthest_point_sampling(x, 4096)
    got, new_xyz = pkg._ext.furthest_point_sampling_gather(x.to(cuda), 4096)
    assert torch.equal(got.cpu(), want)
    centres = new_xyz[:, :300].contiguous()
    for r, ns in ((0.2, 64), (0.4, 32)):
        assert torch.equal(pkg._ext.ball_query(centres, x.to(cuda), r, ns).cpu(),
                           oracle.ext.ball_query(centres.cpu(), x, r, ns)), (r, ns)


@pytest.mark.parametrize("n", [20000, 50000])
def test_fps_batch_of_32_is_scene_independent(pkg, cuda, n):
    """Every scene of a 32-scene batch gets the samples it gets when run alone (cluster scheduling in
    waves and the bucketed kernel included), and samples are unique in-range indices."""
    x = _scene(11, 32, n, 0)[..., :3].contiguous().to(cuda)
    for bucket in (0, 1):
        pkg.set_option("fps_bucket", bucket)
        try:
            full = pkg._ext.furthest_point_sampling(x, 2048)
            for i in (0, 13, 31):
                assert torch.equal(full[i:i + 1], pkg._ext.furthest_point_sampling(x[i:i + 1].contiguous(), 2048)), (bucket, i)
        finally:
            pkg.set_option("fps_bucket", 0)
        assert int(full.min()) >= 0 and int(full.max()) < n
    assert torch.equal(full, pkg._ext.furthest_point_sampling(x, 2048))  # both variants agree


def _fused_vs_generic(pkg, cuda, pc, npoint1):
    from common import fill_state_dict
    bb = pkg.Pointnet2Backbone(input_feature_dim=pc.shape[-1] - 3)
    fill_state_dict(bb)
    bb = bb.to(cuda).eval()
    bb.sa1.npoint = npoint1
    with torch.no_grad():
        fused = bb({"point_clouds": pc})
    assert bb.sa1._fused not in (None, False), "fused path was not taken"
    generic = bb({"point_clouds": pc})   # grad enabled -> op-by-op fp32 path (pointnet2_modules.py)
    for k in ("sa1_inds", "sa2_inds", "fp2_inds", "sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "fp2_xyz"):
        assert torch.equal(fused[k], generic[k]), k
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        a, b = fused[k].float(), generic[k].detach()
        assert torch.isfinite(a).all(), k
        err = (a - b).abs().max().item() / b.abs().max().item()
        assert err < 3e-2, (k, err)
    return fused


def test_backbone_100k_points_npoint_4096(pkg, cuda):
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = _fused_vs_generic(pkg, cuda, _scene(21, 2, 100000, 1).to(cuda), 4096)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    assert out["sa1_inds"].shape == (2, 4096) and out["fp2_features"].shape == (2, 256, 1024)
    assert torch.equal(out["fp2_inds"], out["sa1_inds"][:, :1024])


def test_backbone_multiview_132_channels(pkg, cuda):
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = _fused_vs_generic(pkg, cuda, _scene(31, 2, 40000, 132).to(cuda), 2048)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    assert out["sa1_features"].shape == (2, 128, 2048)
