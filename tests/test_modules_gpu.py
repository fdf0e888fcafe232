"""GPU parity of the module layer (PointnetSAModuleVotes, PointnetFPModule, Pointnet2Backbone)
against the golden fixtures produced by the reference's own Python layers and against the oracle's
functional restatement.  Indices / coordinates exact; fp32 features 1e-4 (generic path, TF32 off);
the fused bf16 tensor-core path is checked in test_fused_gpu.py with its own stated tolerance."""
import numpy as np
import pytest
import torch

from common import cloud, fill_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_convs(pkg):
    """These tests pin the op-by-op SEMANTICS of the module layer at fp32 accuracy (rtol 1e-4 against the
    reference's own layers): TF32 off and the shared MLPs on torch's layers.  The tcgen05 training path has
    its own tests with its own stated tolerance (tests/test_train_gpu.py)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with pkg.mlp_backend("torch"):
        yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _scene():
    pc = cloud(21, 2, 300, c_extra=5, extent=1.0, dup_frac=0.1, origin_pts=3)
    return pc[..., :3].contiguous(), pc[..., 3:].transpose(1, 2).contiguous()


@pytest.mark.parametrize("tag", ["max", "avg", "rbf", "nofeat"])
def test_sa_module_generic_path(pkg, golden, cuda, tag):
    xyz, feats = _scene()
    kw = {"pooling": tag} if tag in ("avg", "rbf") else {}
    m = pkg.PointnetSAModuleVotes(npoint=40, radius=0.45, nsample=12, mlp=[0 if tag == "nofeat" else 5, 16, 16, 32],
                                  use_xyz=True, normalize_xyz=True, **kw)
    fill_state_dict(m)
    m = m.to(cuda).eval()
    with torch.enable_grad():  # grad mode on -> generic op-by-op path
        nx, nf, ii = m(xyz.to(cuda), None if tag == "nofeat" else feats.to(cuda))
    assert (ii.cpu().numpy() == golden[f"sa_{tag}_inds"]).all()
    np.testing.assert_array_equal(nx.cpu().numpy(), golden[f"sa_{tag}_new_xyz"])
    np.testing.assert_allclose(nf.detach().cpu().numpy(), golden[f"sa_{tag}_features"], rtol=1e-4, atol=1e-5)


def test_sa_module_train_mode_and_user_inds(pkg, golden, cuda):
    xyz, feats = _scene()
    m = pkg.PointnetSAModuleVotes(npoint=40, radius=0.45, nsample=12, mlp=[5, 16, 16, 32], use_xyz=True,
                                  normalize_xyz=False)
    fill_state_dict(m)
    m = m.to(cuda).train()
    user_inds = (torch.arange(40, dtype=torch.int32).flip(0).repeat(2, 1).contiguous() * 3).to(cuda)
    nx, nf, ii = m(xyz.to(cuda), feats.to(cuda), user_inds)
    assert torch.equal(ii, user_inds)
    np.testing.assert_array_equal(nx.cpu().numpy(), golden["sa_train_new_xyz"])
    np.testing.assert_allclose(nf.detach().cpu().numpy(), golden["sa_train_features"], rtol=1e-3, atol=1e-4)
    sd = m.state_dict()
    np.testing.assert_allclose(sd["mlp_module.layer0.bn.bn.running_mean"].cpu().numpy(),
                               golden["sa_train_running_mean0"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(sd["mlp_module.layer0.bn.bn.running_var"].cpu().numpy(),
                               golden["sa_train_running_var0"], rtol=1e-4, atol=1e-5)
    nf.sum().backward()  # autograd reaches the conv weights through the grouped gather
    assert m.mlp_module.layer0.conv.weight.grad is not None


@pytest.mark.parametrize("user_inds", [False, True])
def test_sa_module_gradient_reaches_xyz(pkg, cuda, user_inds):
    """d(loss)/d(xyz) through new_xyz (the FPS gather) AND through the recentred grouped coordinates, as in
    the reference where both go through differentiable gather/group ops (pointnet2_modules.py:233-248;
    ProposalModule samples vote_xyz, which carries grad).  Checked against torch's own indexing autograd
    on the same indices."""
    xyz, feats = _scene()
    m = pkg.PointnetSAModuleVotes(npoint=40, radius=0.45, nsample=12, mlp=[5, 16, 16, 32], use_xyz=True,
                                  normalize_xyz=True)
    fill_state_dict(m)
    m = m.to(cuda).eval()   # eval-mode BN keeps the comparison free of batch statistics
    x = xyz.to(cuda).requires_grad_(True)
    f = feats.to(cuda).requires_grad_(True)
    inds_in = (torch.arange(40, dtype=torch.int32).repeat(2, 1) * 5).contiguous().to(cuda) if user_inds else None
    nx, nf, ii = m(x, f, inds_in)
    w_xyz = torch.linspace(-1, 1, nx.numel(), device=cuda).view_as(nx)
    w_f = torch.linspace(0.5, -0.5, nf.numel(), device=cuda).view_as(nf)
    ((nx * w_xyz).sum() + (nf * w_f).sum()).backward()
    assert x.grad is not None and f.grad is not None

    x2 = xyz.to(cuda).requires_grad_(True)
    f2 = feats.to(cuda).requires_grad_(True)
    idx = pkg.pointnet2_utils.ball_query(0.45, 12, xyz.to(cuda), nx.detach()).long()       # (B, 40, 12)
    nx2 = torch.gather(x2, 1, ii.long().unsqueeze(-1).expand(-1, -1, 3))
    b = x2.shape[0]
    ar = torch.arange(b, device=cuda)[:, None, None]
    local = (x2[ar, idx] - nx2.unsqueeze(2)) / 0.45                                          # (B, 40, 12, 3)
    grouped = torch.cat([local.permute(0, 3, 1, 2), f2.transpose(1, 2)[ar, idx].permute(0, 3, 1, 2)], 1)
    nf2 = torch.nn.functional.max_pool2d(m.mlp_module(grouped), kernel_size=[1, 12]).squeeze(-1)
    ((nx2 * w_xyz).sum() + (nf2 * w_f).sum()).backward()
    torch.testing.assert_close(nf, nf2, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(x.grad, x2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(f.grad, f2.grad, rtol=1e-4, atol=1e-5)
    assert x.grad.abs().sum() > 0


def test_fp_module_generic_path(pkg, golden, cuda):
    unknown = cloud(31, 2, 50, extent=1.0)
    known = cloud(32, 2, 20, extent=1.0)
    known[:, 3] = unknown[:, 7]
    uf = cloud(33, 2, 50, c_extra=5)[..., :8].transpose(1, 2).contiguous()
    kf = cloud(34, 2, 20, c_extra=9)[..., :12].transpose(1, 2).contiguous()
    fp = pkg.PointnetFPModule(mlp=[20, 24, 16])
    fill_state_dict(fp)
    fp = fp.to(cuda).eval()
    with torch.enable_grad():
        out = fp(unknown.to(cuda), known.to(cuda), uf.to(cuda), kf.to(cuda))
        out2 = fp(unknown.to(cuda), None, uf.to(cuda), kf[:, :, :1].contiguous().to(cuda))
    np.testing.assert_allclose(out.detach().cpu().numpy(), golden["fp_out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out2.detach().cpu().numpy(), golden["fp_known_none_out"], rtol=1e-4, atol=1e-5)


def test_backbone_generic_path_matches_reference_layers(pkg, golden, cuda):
    bb = pkg.Pointnet2Backbone(input_feature_dim=1)
    fill_state_dict(bb)
    bb = bb.to(cuda).eval()
    scene = cloud(41, 1, 3000, c_extra=1, extent=1.6, dup_frac=0.05, origin_pts=4)
    with torch.enable_grad():
        d = bb({"point_clouds": scene.to(cuda)})
    for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
        assert (d[k].cpu().numpy() == golden["bb_" + k]).all(), k
    np.testing.assert_array_equal(d["sa4_xyz"].cpu().numpy(), golden["bb_sa4_xyz"])
    np.testing.assert_allclose(d["sa1_features"][:, ::4, ::32].detach().cpu().numpy(), golden["bb_sa1_features_sub"],
                               rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(d["fp2_features"][:, ::8, ::16].detach().cpu().numpy(), golden["bb_fp2_features_sub"],
                               rtol=1e-3, atol=1e-4)
    s = d["fp2_features"].double()
    np.testing.assert_allclose([s.sum().item(), s.abs().sum().item()], golden["bb_fp2_features_sum"], rtol=1e-4)


def test_reference_state_dict_round_trip(pkg, cuda):
    """A state_dict saved from one backbone loads into another and reproduces the outputs."""
    a = pkg.Pointnet2Backbone(1)
    fill_state_dict(a)
    b = pkg.Pointnet2Backbone(1)
    missing = b.load_state_dict(a.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
