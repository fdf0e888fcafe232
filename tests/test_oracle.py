"""CPU tests: the oracle against the golden fixtures produced by the reference's own Python layers
(tests/golden/make_golden.py) and against independent restatements of itself."""
import numpy as np
import pytest
import torch

from common import cloud, fill_state_dict


def test_nn_distance_known_answer(oracle, golden):
    """utils/nn_distance.py:95-122, np.random.seed(0): the only known-answer check the reference has."""
    np.random.seed(0)
    pc1 = torch.from_numpy(np.random.random((1, 5, 3)).astype(np.float32))
    pc2 = torch.from_numpy(np.random.random((1, 6, 3)).astype(np.float32))
    d1, i1, d2, i2 = oracle.nn_distance(pc1, pc2)
    np.testing.assert_allclose(d1.numpy()[0], [0.1058, 0.0842, 0.1167, 0.0197, 0.2633], atol=5e-5)
    assert i1.numpy()[0].tolist() == [5, 2, 1, 5, 5]
    d1s, i1s, _, _ = oracle.nn_distance(pc1, pc2, l1smooth=True)
    np.testing.assert_allclose(d1s.numpy()[0], [0.0529, 0.0421, 0.0583, 0.0099, 0.1317], atol=5e-5)
    for tag, kw in (("l2", {}), ("huber", {"l1smooth": True}), ("l1", {"l1": True})):
        r = oracle.nn_distance(pc1, pc2, **kw)
        for got, key in zip(r, ("d1", "i1", "d2", "i2")):
            want = golden[f"nnd_demo_{tag}_{key}"]
            if key[0] == "d":
                np.testing.assert_allclose(got.numpy(), want, rtol=1e-6, atol=1e-7)
            else:
                assert (got.numpy() == want).all()


def test_nn_distance_modes_and_ties(oracle, golden):
    a = cloud(11, 3, 37, extent=1.0)
    b = cloud(12, 3, 23, extent=1.0)
    b[:, 5] = 0.0; b[:, 6] = 0.0
    for tag, kw in (("l2", {}), ("huber", {"l1smooth": True, "delta": 0.15}), ("l1", {"l1": True})):
        r = oracle.nn_distance(a, b, **kw)
        rt = oracle.nn_distance_torch(a, b, **kw)
        for got, got_t, key in zip(r, rt, ("d1", "i1", "d2", "i2")):
            want = golden[f"nnd_b_{tag}_{key}"]
            if key[0] == "d":
                np.testing.assert_allclose(got.numpy(), want, rtol=1e-6, atol=1e-7)
                np.testing.assert_allclose(got_t.numpy(), want, rtol=1e-6, atol=1e-7)
            else:
                assert (got.numpy() == want).all()
                assert (got_t.numpy() == want).all()


def _lattice(k):
    g = torch.stack(torch.meshgrid(*[torch.arange(float(k))] * 3, indexing="ij"), -1)
    return g.reshape(1, -1, 3).contiguous()


@pytest.mark.parametrize("case", ["uniform", "dups", "identical", "lattice", "lattice300", "origin0",
                                  "all_origin", "n1000", "n300", "n3", "npoint1", "npoint_n", "nan"])
def test_fps_key_formulation_matches_thread_tree(oracle, case):
    """The packed-key arg-max (what the CUDA kernel computes) == the literal 512-thread + halving
    tree simulation of sampling_gpu.cu:75-178, on every tie-heavy fixture of SURVEY.md section 4."""
    m = 64
    if case == "uniform":
        x = cloud(1, 2, 5000)[..., :3].contiguous(); m = 256
    elif case == "dups":
        x = cloud(2, 2, 4000, dup_frac=0.5)[..., :3].contiguous(); m = 512
    elif case == "identical":
        x = torch.ones(1, 700, 3) * 0.7
    elif case == "lattice":
        x = _lattice(12); m = 700
    elif case == "lattice300":
        x = _lattice(12)[:, :300].contiguous(); m = 300
    elif case == "origin0":
        x = cloud(3, 1, 2048, origin_pts=40)[..., :3].contiguous(); x[0, 0] = 0.001
    elif case == "all_origin":
        x = (cloud(4, 1, 600)[..., :3] * 0.004).contiguous()
    elif case == "n1000":
        x = cloud(5, 1, 1000)[..., :3].contiguous(); m = 333
    elif case == "n300":
        x = cloud(6, 1, 300)[..., :3].contiguous(); m = 150
    elif case == "n3":
        x = cloud(7, 1, 3)[..., :3].contiguous(); m = 3
    elif case == "npoint1":
        x = cloud(8, 1, 100)[..., :3].contiguous(); m = 1
    elif case == "npoint_n":
        x = cloud(9, 1, 257)[..., :3].contiguous(); m = 257
    else:
        x = cloud(10, 1, 900)[..., :3].contiguous()
        x[0, 17, 1] = float("nan"); x[0, 400] = float("inf"); m = 40
    a = oracle.ext.furthest_point_sampling(x, m)
    b = oracle.furthest_point_sampling_keyed(x, m)
    assert (a == b).all()
    assert a[:, 0].eq(0).all()


def test_opt_n_threads(oracle):
    # cuda_utils.h:20-24 evaluated numerically: 512 for every size >= 512 on the path
    for n, want in ((40000, 512), (2048, 512), (1024, 512), (512, 512), (300, 256), (3, 2), (1, 1)):
        assert oracle.opt_n_threads(n) == want


def test_ball_query_semantics(oracle):
    xyz = torch.tensor([[[0.0, 0, 0], [0.1, 0, 0], [0.2, 0, 0], [5.0, 5, 5], [0.05, 0, 0], [0.3, 0, 0]]])
    centres = torch.tensor([[[0.0, 0, 0], [9.0, 9, 9], [0.25, 0, 0]]])
    idx = oracle.ext.ball_query(centres, xyz, 0.2, 4)
    # centre 0: hits 0,1,4 (0.2 is NOT < 0.2*0.2 in fp32? d2 = 0.04 vs r2 = fl(0.2f*0.2f)) -> padded with first hit
    r2 = np.float32(0.2) * np.float32(0.2)
    d2 = np.float32(0.2) * np.float32(0.2)
    expect0 = [0, 1, 2, 4] if d2 < r2 else [0, 1, 4, 0]
    assert idx[0, 0].tolist() == expect0
    assert idx[0, 1].tolist() == [0, 0, 0, 0]          # empty ball: zeros
    assert idx[0, 2].tolist() == [1, 2, 5, 1]          # 3 hits then first-hit padding
    idx2 = oracle.ext.ball_query(centres, xyz, 10.0, 3)  # more hits than nsample: first three
    assert idx2[0, 0].tolist() == [0, 1, 2]


def test_three_nn_semantics(oracle):
    known = torch.tensor([[[0.0, 0, 0], [1.0, 0, 0], [1.0, 0, 0], [2.0, 0, 0]]])
    unknown = torch.tensor([[[1.0, 0, 0], [0.0, 0, 0]]])
    d2, idx = oracle.ext.three_nn(unknown, known)
    assert idx[0, 0].tolist() == [1, 2, 0]             # equal distances: earlier index first
    assert d2[0, 0].tolist() == [0.0, 0.0, 1.0]
    d2b, idxb = oracle.ext.three_nn(unknown, known[:, :2].contiguous())  # m = 2: slot 3 unfilled
    assert torch.isinf(d2b[0, 0, 2]) and idxb[0, 0, 2] == 0


def test_query_and_group_matches_reference_layer(oracle, golden):
    pc = cloud(21, 2, 300, c_extra=5, extent=1.0, dup_frac=0.1, origin_pts=3)
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
    inds = oracle.ext.furthest_point_sampling(xyz, 40)
    new_xyz = oracle.ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    nf, gx, _ = oracle.query_and_group(xyz, new_xyz, feats, 0.45, 12, normalize_xyz=True)
    np.testing.assert_array_equal(nf.numpy(), golden["qg_new_features"])
    np.testing.assert_array_equal(gx.numpy(), golden["qg_grouped_xyz"])


def _sa_state(pkg, mlp, **kw):
    m = pkg.PointnetSAModuleVotes(npoint=40, radius=0.45, nsample=12, mlp=list(mlp), use_xyz=True, **kw)
    return fill_state_dict(m)


@pytest.mark.parametrize("tag", ["max", "avg", "rbf", "nofeat"])
def test_sa_module_matches_reference_layer(oracle, golden, pkg, tag):
    pc = cloud(21, 2, 300, c_extra=5, extent=1.0, dup_frac=0.1, origin_pts=3)
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous() if tag != "nofeat" else None
    sd = _sa_state(pkg, [5 if tag != "nofeat" else 0, 16, 16, 32], normalize_xyz=True)
    pooling = tag if tag in ("avg", "rbf") else "max"
    nx, nf, ii = oracle.sa_module(xyz, feats, sd, "", 40, 0.45, 12, normalize_xyz=True, pooling=pooling)
    assert (ii.numpy() == golden[f"sa_{tag}_inds"]).all()
    np.testing.assert_array_equal(nx.numpy(), golden[f"sa_{tag}_new_xyz"])
    np.testing.assert_allclose(nf.numpy(), golden[f"sa_{tag}_features"], rtol=1e-5, atol=1e-6)


def test_sa_module_train_mode_matches_reference_layer(oracle, golden, pkg):
    pc = cloud(21, 2, 300, c_extra=5, extent=1.0, dup_frac=0.1, origin_pts=3)
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
    sd = _sa_state(pkg, [5, 16, 16, 32], normalize_xyz=False)
    user_inds = torch.arange(40, dtype=torch.int32).flip(0).repeat(2, 1).contiguous() * 3
    nx, nf, _ = oracle.sa_module(xyz, feats, sd, "", 40, 0.45, 12, normalize_xyz=False, inds=user_inds,
                                 training=True)
    np.testing.assert_array_equal(nx.numpy(), golden["sa_train_new_xyz"])
    np.testing.assert_allclose(nf.numpy(), golden["sa_train_features"], rtol=1e-4, atol=1e-5)


def test_fp_module_matches_reference_layer(oracle, golden, pkg):
    unknown = cloud(31, 2, 50, extent=1.0)
    known = cloud(32, 2, 20, extent=1.0)
    known[:, 3] = unknown[:, 7]
    uf = cloud(33, 2, 50, c_extra=5)[..., :8].transpose(1, 2).contiguous()
    kf = cloud(34, 2, 20, c_extra=9)[..., :12].transpose(1, 2).contiguous()
    sd = fill_state_dict(pkg.PointnetFPModule(mlp=[20, 24, 16]))
    out = oracle.fp_module(unknown, known, uf, kf, sd, "")
    np.testing.assert_allclose(out.numpy(), golden["fp_out"], rtol=1e-5, atol=1e-6)
    out2 = oracle.fp_module(unknown, None, uf, kf[:, :, :1].contiguous(), sd, "")
    np.testing.assert_allclose(out2.numpy(), golden["fp_known_none_out"], rtol=1e-5, atol=1e-6)


def test_backbone_matches_reference_layer(oracle, golden, pkg):
    """Whole Pointnet2Backbone (backbone_module.py:76-135) on a 3000-point scene with duplicates and
    near-origin points: indices exact, features to fp32 round-off."""
    sd = fill_state_dict(pkg.Pointnet2Backbone(input_feature_dim=1))
    scene = cloud(41, 1, 3000, c_extra=1, extent=1.6, dup_frac=0.05, origin_pts=4)
    d = oracle.backbone(scene, sd)
    for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
        assert (d[k].numpy() == golden["bb_" + k]).all(), k
    np.testing.assert_array_equal(d["sa4_xyz"].numpy(), golden["bb_sa4_xyz"])
    np.testing.assert_allclose(d["sa1_features"][:, ::4, ::32].numpy(), golden["bb_sa1_features_sub"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(d["sa4_features"][:, ::8, ::8].numpy(), golden["bb_sa4_features_sub"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(d["fp2_features"][:, ::8, ::16].numpy(), golden["bb_fp2_features_sub"], rtol=1e-4, atol=1e-5)
    s = d["fp2_features"].double()
    np.testing.assert_allclose([s.sum().item(), s.abs().sum().item()], golden["bb_fp2_features_sum"], rtol=1e-5)
