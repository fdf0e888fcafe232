"""GPU tests of the tensor-core (tcgen05) path.

Numerics contract of the fused path: operands of every tensor-core contraction (activations and
BatchNorm-folded weights) are rounded to bf16 (RNE), products are exact and accumulate in fp32.
`mixed_reference` below restates exactly that in torch; the kernels must agree with it to
rtol 1e-3 (differences: fp32 summation order, and rare one-ulp bf16 flips of intermediate
activations).  Against the plain fp32 reference the same outputs agree to ~1e-2 of the tensor's
max magnitude (bf16 has 8 mantissa bits); both tolerances are asserted."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (1000, 128, 100), (4096, 256, 512), (130, 16, 64), (77, 256, 259),
                                   (300, 128, 128)])
@pytest.mark.parametrize("out", ["f32", "bf16", "cm"])
def test_linear_bf16(pkg, cuda, m, n, k, out):
    import importlib
    tcm = importlib.import_module("3dvlp_b200.tensorcore")
    g = torch.Generator().manual_seed(m * 7 + n)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(cuda)
    bias = torch.randn(n, generator=g).to(cuda)
    x = torch.randn(m, k, generator=g).to(cuda)
    pw = tcm.PackedWeight(w)
    xp = torch.zeros(m, pw.kpad, device=cuda, dtype=torch.bfloat16)
    xp[:, :k] = x.to(torch.bfloat16)
    rpb = m // 2 if (out == "cm" and m % 2 == 0) else m
    y = tcm.linear(xp, pw, bias, relu=True, out=out, rows_per_batch=rpb)
    assert tcm.poll_device_error(cuda) == 0
    want = torch.relu(bf16r(x).double() @ bf16r(w).double().t() + bias.double()).float()
    if out == "cm":
        y = y.permute(0, 2, 1).reshape(m, n)
    if out == "bf16":
        torch.testing.assert_close(y.float(), want, rtol=1e-2, atol=1e-2)
    else:
        torch.testing.assert_close(y, want, rtol=1e-3, atol=1e-4)


def test_pack_rows(pkg, cuda):
    import importlib
    tcm = importlib.import_module("3dvlp_b200.tensorcore")
    f = torch.randn(3, 131, 777, device=cuda)
    p = tcm.pack_rows(f)
    assert p.shape == (3 * 777, 192)
    want = torch.zeros(3, 777, 192, device=cuda)
    want[:, :, :131] = f.transpose(1, 2)
    assert torch.equal(p.float().view(3, 777, 192), bf16r(want))
    # a width that is not a multiple of eight takes the scalar-store kernel; exact tile multiples too
    p2 = tcm.pack_rows(f, kpad=132)
    assert torch.equal(p2.float().view(3, 777, 132), bf16r(want[:, :, :132]))
    g = torch.randn(2, 128, 2048, device=cuda)
    assert torch.equal(tcm.pack_rows(g).float().view(2, 2048, 128), bf16r(g.transpose(1, 2)))
    h = torch.randn(1, 5, 33, device=cuda)
    p3 = tcm.pack_rows(h)
    assert p3.shape == (33, 64) and torch.equal(p3.float()[:, :5], bf16r(h[0].t())) and p3.float()[:, 5:].eq(0).all()


SA_CASES = {
    # name: (B, N, C_in, npoint, radius, ns, mlp)
    "sa1_like": (2, 3000, 1, 256, 0.3, 64, [64, 64, 128]),
    "sa1_rgbh": (1, 2000, 4, 128, 0.3, 64, [64, 64, 128]),
    "sa1_nofeat": (2, 1500, 0, 100, 0.35, 32, [64, 64, 128]),
    "sa2_like": (2, 2048, 128, 512, 0.4, 32, [128, 128, 256]),
    "sa3_like": (3, 1024, 256, 250, 0.8, 16, [128, 128, 256]),
    "voteagg_like": (2, 1024, 256, 256, 0.3, 16, [128, 128, 128]),
    "wide_input": (1, 2500, 132, 200, 0.3, 64, [64, 64, 128]),
    # the fused kernel skips ball-query padding rows: nearly empty balls (1-3 distinct hits, hundreds of
    # centres per tile), completely full balls (no row to skip) and a ragged tail (centres straddle tiles)
    "sparse_balls": (2, 3000, 1, 300, 0.03, 64, [64, 64, 128]),
    "full_balls": (2, 1200, 1, 130, 5.0, 32, [64, 64, 128]),
    "sparse_balls_wide": (2, 2048, 128, 517, 0.05, 32, [128, 128, 256]),
    "ns128": (1, 900, 1, 37, 0.6, 128, [64, 64, 128]),          # a full ball is exactly one tile
    "tiny": (1, 40, 0, 3, 0.5, 16, [64, 64, 128]),               # fewer rows than one tile
}


@pytest.mark.parametrize("name", sorted(SA_CASES))
def test_sa_fused_matches_mixed_precision_contract(pkg, oracle, cuda, name):
    from common import cloud, fill_state_dict
    b, n, c_in, npoint, radius, ns, mlp = SA_CASES[name]
    pc = cloud(100 + n, b, n, c_extra=c_in, extent=1.2, dup_frac=0.05, origin_pts=2)
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous() if c_in else None
    if feats is not None:
        feats = feats * 2.0 - 0.7
    m = pkg.PointnetSAModuleVotes(npoint=npoint, radius=radius, nsample=ns, mlp=[c_in] + mlp, use_xyz=True,
                                  normalize_xyz=True)
    sd = fill_state_dict(m)
    m = m.to(cuda).eval()
    with torch.no_grad():
        nx, nf, ii = m(xyz.to(cuda), None if feats is None else feats.to(cuda))
    import importlib
    tcm = importlib.import_module("3dvlp_b200.tensorcore")
    assert tcm.poll_device_error(cuda) == 0
    assert m._fused not in (None, False), "fused path was not taken"
    wx, wf, wi = oracle.sa_module_mixed(xyz, feats, sd, "", npoint, radius, ns, normalize_xyz=True)
    assert torch.equal(ii.cpu(), wi)
    assert torch.equal(nx.cpu(), wx)
    scale = wf.abs().max().item()
    # rtol 1e-3; the absolute floor (4e-3 of the tensor's max) covers the rare one-ulp bf16 flips of
    # intermediate activations (2^-8 relative on one of K terms, |w| ~ 0.2) caused by fp32 summation
    # order; fewer than 1 element in 10^3 may exceed 1e-3 at all (next assert)
    torch.testing.assert_close(nf.cpu(), wf, rtol=1e-3, atol=4e-3 * scale)
    assert ((nf.cpu() - wf).abs() > 1e-3 * scale + 1e-3 * wf.abs()).float().mean().item() < 1e-3
    # and against the plain fp32 reference semantics: bf16 operand rounding only
    _, ff, _ = oracle.sa_module(xyz, feats, sd, "", npoint, radius, ns, normalize_xyz=True)
    err = (nf.cpu() - ff).abs().max().item() / ff.abs().max().item()
    assert err < 2e-2, err
    # the bf16 point-major twin handed to the next layer
    rows = nf._pn2_rows[1].float().view(b, npoint, -1).transpose(1, 2)
    torch.testing.assert_close(rows.cpu(), nf.cpu(), rtol=1e-2, atol=1e-2 * scale)
    # an in-place edit of the fp32 tensor invalidates the twin (it is keyed on the tensor version)
    import importlib
    fused = importlib.import_module("3dvlp_b200.fused")
    assert fused.cached_rows(nf) is not None
    nf.mul_(2.0)
    assert fused.cached_rows(nf) is None


def test_sa_fused_user_inds_and_unnormalised(pkg, oracle, cuda):
    from common import cloud, fill_state_dict
    pc = cloud(333, 2, 1200, c_extra=1, extent=1.0)
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
    m = pkg.PointnetSAModuleVotes(npoint=64, radius=0.4, nsample=16, mlp=[1, 64, 64, 128], use_xyz=True,
                                  normalize_xyz=False)
    sd = fill_state_dict(m)
    m = m.to(cuda).eval()
    inds = (torch.arange(64, dtype=torch.int32) * 7).repeat(2, 1).contiguous()
    with torch.no_grad():
        nx, nf, ii = m(xyz.to(cuda), feats.to(cuda), inds.to(cuda))
    wx, wf, _ = oracle.sa_module_mixed(xyz, feats, sd, "", 64, 0.4, 16, normalize_xyz=False, inds=inds)
    assert torch.equal(nx.cpu(), wx) and torch.equal(ii.cpu(), inds)
    torch.testing.assert_close(nf.cpu(), wf, rtol=1e-3, atol=2e-3 * wf.abs().max().item())


@pytest.mark.parametrize("n,m,c1,c2,mlp", [(512, 256, 256, 256, [256, 256]), (1024, 512, 256, 256, [256, 256]),
                                          (300, 70, 19, 45, [128, 64]), (50, 2, 0, 12, [64])])
def test_fp_fused_matches_mixed_precision_contract(pkg, oracle, cuda, n, m, c1, c2, mlp):
    from common import cloud, fill_state_dict
    unknown = cloud(31 + n, 2, n, extent=1.0)[..., :3].contiguous()
    known = cloud(32 + m, 2, m, extent=1.0)[..., :3].contiguous()
    if m > 3:
        known[:, 3] = unknown[:, 7]
    g = torch.Generator().manual_seed(n + m)
    uf = torch.randn(2, c1, n, generator=g) if c1 else None
    kf = torch.randn(2, c2, m, generator=g)
    fp = pkg.PointnetFPModule(mlp=[c1 + c2] + mlp)
    sd = fill_state_dict(fp)
    fp = fp.to(cuda).eval()
    with torch.no_grad():
        got = fp(unknown.to(cuda), known.to(cuda), None if uf is None else uf.to(cuda), kf.to(cuda))
    if c1 == 0:
        assert fp._fused is None  # unknow_feats=None takes the generic path by design
        return
    assert fp._fused not in (None, False), "fused path was not taken"
    want = oracle.fp_module_mixed(unknown, known, uf, kf, sd, "", n_layers=len(mlp))
    scale = want.abs().max().item()
    torch.testing.assert_close(got.cpu(), want, rtol=1e-3, atol=4e-3 * scale)
    ref = oracle.fp_module(unknown, known, uf, kf, sd, "", n_layers=len(mlp))
    assert (got.cpu() - ref).abs().max().item() / ref.abs().max().item() < 2e-2


def test_backbone_fused_vs_oracle(pkg, oracle, cuda):
    """Whole backbone through the fused path on a 5000-point scene: indices/coordinates bit-exact,
    features within the mixed-precision tolerance of the fp32 reference semantics."""
    from common import cloud, fill_state_dict
    bb = pkg.Pointnet2Backbone(input_feature_dim=1)
    sd = fill_state_dict(bb)
    bb = bb.to(cuda).eval()
    scene = cloud(41, 2, 5000, c_extra=1, extent=1.6, dup_frac=0.05, origin_pts=4)
    with torch.no_grad():
        d = bb({"point_clouds": scene.to(cuda)})
    want = oracle.backbone(scene, sd)
    for k in ("sa1_inds", "sa2_inds", "fp2_inds", "sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "fp2_xyz"):
        assert torch.equal(d[k].cpu(), want[k]), k
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        err = (d[k].cpu() - want[k]).abs().max().item() / want[k].abs().max().item()
        assert err < 3e-2, (k, err)
    assert d["fp2_features"].shape == (2, 256, 1024) and d["fp2_inds"].dtype == torch.int32


def test_graphed_backbone_replays_identically(pkg, cuda):
    """CUDA-graph replay == eager forward, on fresh inputs copied into the static buffer."""
    import importlib
    from common import fill_state_dict
    syn = importlib.import_module("3dvlp_b200.synthetic")
    bb = pkg.Pointnet2Backbone(input_feature_dim=1)
    fill_state_dict(bb)
    bb = bb.to(cuda).eval()
    a = syn.make_batch(0, 2, 9000, 1).to(cuda)   # 9000 points: cluster FPS + grid ball query
    b2 = syn.make_batch(5, 2, 9000, 1).to(cuda)
    g = pkg.GraphedBackbone(bb, a)
    assert g.graph is not None, g.capture_error
    for pc in (b2, a, b2):
        out = g(pc)
        torch.cuda.synchronize()
        got = {k: out[k].clone() for k in ("sa1_inds", "sa2_inds", "fp2_inds", "fp2_xyz", "fp2_features", "sa4_features")}
        with torch.no_grad():
            want = bb({"point_clouds": pc.clone()})
        for k, v in got.items():
            assert torch.equal(v, want[k]), k


def test_graphed_backbone_packed_outputs(pkg, cuda):
    """GraphedBackbone(pack=...) places the outputs a serving loop reads back in ONE contiguous buffer (the last
    linear kernel writes fp2_features straight into it): one device->host copy returns the same tensors."""
    import importlib
    from common import fill_state_dict
    syn = importlib.import_module("3dvlp_b200.synthetic")
    bb = pkg.Pointnet2Backbone(input_feature_dim=1)
    fill_state_dict(bb)
    bb = bb.to(cuda).eval()
    a = syn.make_batch(0, 2, 9000, 1).to(cuda)
    keys = ("fp2_features", "fp2_xyz", "fp2_inds")
    g = pkg.GraphedBackbone(bb, a, pack=keys, priority=-1)
    assert g.graph is not None, g.capture_error
    host = torch.empty(g.packed.numel(), dtype=torch.uint8).pin_memory()
    for pc in (syn.make_batch(5, 2, 9000, 1).to(cuda), a):
        out = g(pc)
        host.copy_(g.packed, non_blocking=True)
        torch.cuda.synchronize()
        with torch.no_grad():
            want = bb({"point_clouds": pc.clone()})
        views = g.packed_views(host)
        for k in keys:
            assert views[k].dtype == want[k].dtype and views[k].shape == want[k].shape
            assert torch.equal(views[k], want[k].cpu()), k
            assert torch.equal(out[k], want[k]), k
    assert out["fp2_features"].data_ptr() == g.packed.data_ptr()  # written in place, no copy node


@pytest.mark.parametrize("name", ["sa1_like", "sa2_like", "sa3_like", "sparse_balls_wide", "tiny"])
def test_sa_tiles_per_group_option_is_bit_identical(pkg, cuda, name):
    """pn2_set_option("sa_tiles_per_group", k) only changes how many persistent CTAs share the tiles
    (throughput mode); tiles are claimed dynamically, the output must not move by a bit."""
    from common import cloud, fill_state_dict
    b, n, c_in, npoint, radius, ns, mlp = SA_CASES[name]
    pc = cloud(100 + n, b, n, c_extra=c_in, extent=1.2, dup_frac=0.05, origin_pts=2)
    xyz = pc[..., :3].contiguous().to(cuda)
    feats = (pc[..., 3:].transpose(1, 2).contiguous() * 2.0 - 0.7).to(cuda) if c_in else None
    m = pkg.PointnetSAModuleVotes(npoint=npoint, radius=radius, nsample=ns, mlp=[c_in] + mlp, use_xyz=True,
                                  normalize_xyz=True)
    fill_state_dict(m)
    m = m.to(cuda).eval()
    outs = []
    try:
        for k in (0, 1, 16, 1000):
            pkg.set_option("sa_tiles_per_group", k)
            assert pkg.get_option("sa_tiles_per_group") == k
            with torch.no_grad():
                _, nf, _ = m(xyz, feats)
            outs.append(nf.clone())
    finally:
        pkg.set_option("sa_tiles_per_group", 0)
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


@pytest.mark.parametrize("c_in", [0, 1, 4])
def test_sa_four_group_cta_equals_single_group_ctas(pkg, cuda, c_in):
    """SA1-sized weights run four tile groups in ONE CTA per SM once a call holds >= 4 tiles per SM (the prologue
    is paid once per SM), and four single-group CTAs per SM below that (csrc/sa_fused.cu, pn2_sa_group_mlp).
    The batched call (6 x 256 x 64 / 128 = 768 tiles) takes the first layout, each scene alone (128 tiles) the
    second: same bits."""
    from common import cloud, fill_state_dict
    b, n, npoint, ns = 6, 4000, 256, 64
    pc = cloud(77 + c_in, b, n, c_extra=c_in, extent=1.5, dup_frac=0.05, origin_pts=2)
    xyz = pc[..., :3].contiguous().to(cuda)
    feats = (pc[..., 3:].transpose(1, 2).contiguous() * 2.0 - 0.7).to(cuda) if c_in else None
    m = pkg.PointnetSAModuleVotes(npoint=npoint, radius=0.3, nsample=ns, mlp=[c_in, 64, 64, 128], use_xyz=True,
                                  normalize_xyz=True)
    fill_state_dict(m)
    m = m.to(cuda).eval()
    with torch.no_grad():
        _, nf, inds = m(xyz, feats)
        for i in range(b):
            _, nf_i, inds_i = m(xyz[i:i + 1].contiguous(), None if feats is None else feats[i:i + 1].contiguous())
            assert torch.equal(inds_i[0], inds[i])
            assert torch.equal(nf_i[0], nf[i]), i


def test_throughput_graph_equals_latency_graph(pkg, cuda):
    """GraphedBackbone(throughput=True) (bucketed FPS, fewer fused-SA CTAs) == the default graph."""
    import importlib
    from common import fill_state_dict
    syn = importlib.import_module("3dvlp_b200.synthetic")
    bb = pkg.Pointnet2Backbone(input_feature_dim=1)
    fill_state_dict(bb)
    bb = bb.to(cuda).eval()
    a = syn.make_batch(3, 2, 20000, 1).to(cuda)
    g0 = pkg.GraphedBackbone(bb, a)
    g1 = pkg.GraphedBackbone(bb, a, throughput=True)
    assert g0.graph is not None and g1.graph is not None
    assert pkg.get_option("sa_tiles_per_group") == 0 and pkg.get_option("fps_bucket") == 0
    o0 = {k: v.clone() for k, v in g0(a).items() if torch.is_tensor(v)}
    o1 = g1(a)
    torch.cuda.synchronize()
    for k, v in o0.items():
        assert torch.equal(v, o1[k]), k


def test_sa_module_beyond_fused_limits_takes_generic_path(pkg, cuda):
    """More than 64 scenes per call (the fused kernel's row-plan limit) silently uses the op-by-op path:
    same indices, features within the mixed-precision tolerance of the fused result on the two halves."""
    from common import cloud, fill_state_dict
    pc = cloud(77, 70, 600, c_extra=1, extent=1.0)
    xyz = pc[..., :3].contiguous().to(cuda)
    feats = pc[..., 3:].transpose(1, 2).contiguous().to(cuda)
    m = pkg.PointnetSAModuleVotes(npoint=64, radius=0.3, nsample=16, mlp=[1, 64, 64, 128], use_xyz=True, normalize_xyz=True)
    fill_state_dict(m)
    m = m.to(cuda).eval()
    with torch.no_grad():
        nx, nf, ii = m(xyz, feats)
        halves = [m(xyz[s].contiguous(), feats[s].contiguous()) for s in (slice(0, 35), slice(35, 70))]
    assert torch.equal(ii, torch.cat([h[2] for h in halves])) and torch.equal(nx, torch.cat([h[0] for h in halves]))
    want = torch.cat([h[1] for h in halves])
    assert (nf - want).abs().max().item() / want.abs().max().item() < 3e-2


@pytest.mark.parametrize("n,m", [(512, 256), (1024, 512), (50, 3), (50, 2), (50, 1), (777, 1500)])
def test_three_nn_weights_kernel(pkg, oracle, cuda, n, m):
    """Warp-per-point neighbour search + weights (run ahead of the FP layers on the sampling stream):
    same neighbours, order and ties as three_nn, weights as PointnetFPModule computes them."""
    import importlib
    from common import cloud
    fused = importlib.import_module("3dvlp_b200.fused")
    u = cloud(160 + n, 2, n, extent=1.0)[..., :3].contiguous()
    k = cloud(161 + m, 2, m, extent=1.0)[..., :3].contiguous()
    if m > 2:
        k[:, 2] = k[:, 1]            # duplicate known points: equal distances, earlier index wins
        k[:, 0] = u[:, 5]            # unknown == known: d = 0
    d2w, iw = oracle.ext.three_nn(u, k)
    idx, w = fused.fp_neighbours(u.to(cuda), k.to(cuda))
    assert torch.equal(idx.cpu(), iw)
    recip = 1.0 / (torch.sqrt(d2w) + 1e-8)      # pointnet2_modules.py:395-397
    want = recip / recip.sum(dim=2, keepdim=True)
    torch.testing.assert_close(w.cpu(), want, rtol=2e-6, atol=1e-7)
