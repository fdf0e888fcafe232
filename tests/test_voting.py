"""VotingModule / VoteNetFrontEnd: CPU checks against the golden produced by
the reference's own VotingModule, GPU checks of the tensor-core fast path."""
import numpy as np
import pytest
import torch

from common import cloud, fill_state_dict, fill_state_dict_conv1d


def _inputs():
    sxyz = cloud(51, 2, 96, extent=1.0)[..., :3].contiguous()
    sfeat = cloud(52, 2, 96, c_extra=61)[..., :64].transpose(1, 2).contiguous()
    return sxyz, sfeat


def test_voting_module_generic_path_matches_reference(pkg, golden):
    """voting_module.py:38-59 run by the reference itself (tests/golden/make_golden.py)."""
    vm = pkg.VotingModule(1, 64)
    fill_state_dict_conv1d(vm)
    vm.eval()
    sxyz, sfeat = _inputs()
    with torch.no_grad():
        vx, vf = vm(sxyz, sfeat)
    np.testing.assert_allclose(vx.numpy(), golden["vote_xyz"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(vf.numpy(), golden["vote_features"], rtol=1e-5, atol=1e-6)


def test_voting_parameter_layout(pkg):
    vm = pkg.VotingModule(1, 256)
    assert sum(p.numel() for p in vm.parameters()) == 199171
    assert set(k.split(".")[0] for k in vm.state_dict()) == {"conv1", "conv2", "conv3", "bn1", "bn2"}
    fe = pkg.VoteNetFrontEnd(input_feature_dim=1)
    keys = fe.state_dict().keys()
    assert "backbone_net.sa1.mlp_module.layer0.conv.weight" in keys
    assert "vgen.conv3.weight" in keys
    assert "proposal.vote_aggregation.mlp_module.layer2.bn.bn.running_var" in keys


@pytest.mark.gpu
def test_voting_module_fused_path(pkg, golden, cuda):
    vm = pkg.VotingModule(1, 64)
    fill_state_dict_conv1d(vm)
    vm = vm.to(cuda).eval()
    sxyz, sfeat = _inputs()
    with torch.no_grad():
        vx, vf = vm(sxyz.to(cuda), sfeat.to(cuda))
    assert vm._fused is not None
    # bf16-operand contract: compare with the fp32 reference at bf16 precision
    scale = float(np.abs(golden["vote_features"]).max())
    np.testing.assert_allclose(vx.cpu().numpy(), golden["vote_xyz"], rtol=0, atol=2e-2)
    np.testing.assert_allclose(vf.cpu().numpy(), golden["vote_features"], rtol=0, atol=2e-2 * scale)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    with torch.enable_grad():  # generic path on the GPU, fp32
        gx, gf = vm(sxyz.to(cuda), sfeat.to(cuda))
    torch.backends.cudnn.allow_tf32 = old
    np.testing.assert_allclose(gx.detach().cpu().numpy(), golden["vote_xyz"], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_votenet_front_end(pkg, oracle, cuda):
    """backbone -> votes -> L2 norm -> vote aggregation: fused path vs the op-by-op path."""
    fe = pkg.VoteNetFrontEnd(input_feature_dim=1)
    fill_state_dict(fe.backbone_net)
    fill_state_dict_conv1d(fe.vgen)
    fill_state_dict(fe.proposal)
    fe = fe.to(cuda).eval()
    scene = cloud(41, 2, 5000, c_extra=1, extent=1.6, dup_frac=0.05, origin_pts=4).to(cuda)
    with torch.no_grad():
        d = fe({"point_clouds": scene})
    with torch.enable_grad():
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        w = fe({"point_clouds": scene.clone()})
    assert d["aggregated_vote_features"].shape == (2, 256, 128)
    assert d["aggregated_vote_inds"].dtype == torch.int32 and d["aggregated_vote_inds"].shape == (2, 256)
    assert torch.equal(d["seed_inds"], w["seed_inds"]) and torch.equal(d["seed_xyz"], w["seed_xyz"])
    for k in ("vote_xyz", "vote_features"):
        err = (d[k] - w[k].detach()).abs().max().item() / w[k].detach().abs().max().item()
        assert err < 5e-2, (k, err)
    # vote aggregation on IDENTICAL votes (FPS over slightly different vote_xyz may legitimately pick
    # different cluster centres, so the end-to-end tensors are not comparable element by element)
    vx, vf = w["vote_xyz"].detach().contiguous(), w["vote_features"].detach().contiguous()
    with torch.no_grad():
        ax, af, ai = fe.proposal.vote_aggregation(vx, vf)
    with torch.enable_grad():
        bx, bf, bi = fe.proposal.vote_aggregation(vx, vf)
    assert torch.equal(ai, bi) and torch.equal(ax, bx)
    err = (af - bf.detach()).abs().max().item() / bf.detach().abs().max().item()
    assert err < 3e-2, err
    torch.testing.assert_close(d["vote_features"].norm(dim=1), torch.ones(2, 1024, device=cuda), rtol=1e-4, atol=1e-4)
