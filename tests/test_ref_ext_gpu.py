"""GPU tests that pin the CPU oracle (and the product) against the UNMODIFIED reference CUDA
extension, compiled for sm_100a by oracle/build_ref.py into oracle/_ref/pointnet2_ref_ext.so.
Skipped when the binary is absent (it is git-ignored; gpurun ships it with the snapshot)."""
import importlib.util
import os

import pytest
import torch

from common import cloud

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "pointnet2_ref_ext.so")


@pytest.fixture(scope="module")
def ref_ext(cuda):
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/pointnet2_ref_ext.so not built (python oracle/build_ref.py)")
    spec = importlib.util.spec_from_file_location("pointnet2_ref_ext", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _lattice(k):
    g = torch.stack(torch.meshgrid(*[torch.arange(float(k))] * 3, indexing="ij"), -1)
    return g.reshape(1, -1, 3).contiguous()


@pytest.mark.parametrize("case", ["uniform", "dups", "lattice", "origin", "n300", "n1000", "scene40k"])
def test_fps_reference_ext_vs_oracle_vs_product(pkg, oracle, ref_ext, cuda, case):
    if case == "uniform":
        x, m = cloud(1, 2, 5000)[..., :3].contiguous(), 256
    elif case == "dups":
        x, m = cloud(2, 2, 4000, dup_frac=0.5)[..., :3].contiguous(), 512
    elif case == "lattice":
        x, m = _lattice(12), 700
    elif case == "origin":
        x, m = cloud(3, 1, 2048, origin_pts=40)[..., :3].contiguous(), 64
    elif case == "n300":
        x, m = cloud(6, 1, 300)[..., :3].contiguous(), 150
    elif case == "n1000":
        x, m = cloud(5, 1, 1000)[..., :3].contiguous(), 333
    else:
        import importlib
        syn = importlib.import_module("3dvlp_b200.synthetic")
        x, m = syn.make_batch(2, 2, 40000, 1)[..., :3].contiguous(), 2048
    ref = ref_ext.furthest_point_sampling(x.to(cuda), m).cpu()
    assert torch.equal(ref, oracle.ext.furthest_point_sampling(x, m)), "oracle != reference ext"
    assert torch.equal(ref, pkg._ext.furthest_point_sampling(x.to(cuda), m).cpu()), "product != reference ext"


def test_search_ops_reference_ext_vs_oracle_vs_product(pkg, oracle, ref_ext, cuda):
    x = cloud(90, 2, 6000, extent=1.0, dup_frac=0.1)[..., :3].contiguous()
    inds = oracle.ext.furthest_point_sampling(x, 256)
    c = torch.gather(x, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    for r, ns in ((0.2, 64), (0.4, 32), (0.8, 16), (1.2, 16), (0.3, 16)):
        ref = ref_ext.ball_query(c.to(cuda), x.to(cuda), r, ns).cpu()
        assert torch.equal(ref, oracle.ext.ball_query(c, x, r, ns)), ("oracle", r)
        assert torch.equal(ref, pkg._ext.ball_query(c.to(cuda), x.to(cuda), r, ns).cpu()), ("product", r)
    u = x[:, :1024].contiguous(); k = c
    d2r, ir = ref_ext.three_nn(u.to(cuda), k.to(cuda))
    d2o, io = oracle.ext.three_nn(u, k)
    d2p, ip = pkg._ext.three_nn(u.to(cuda), k.to(cuda))
    assert torch.equal(ir.cpu(), io) and torch.equal(d2r.cpu(), d2o)
    assert torch.equal(ir, ip) and torch.equal(d2r, d2p)


def test_index_ops_reference_ext_vs_oracle_vs_product(pkg, oracle, ref_ext, cuda):
    g = torch.Generator().manual_seed(5)
    pts = torch.randn(2, 19, 700, generator=g)
    idx = torch.randint(0, 700, (2, 50, 8), generator=g, dtype=torch.int32)
    ref = ref_ext.group_points(pts.to(cuda), idx.to(cuda)).cpu()
    assert torch.equal(ref, oracle.ext.group_points(pts, idx))
    assert torch.equal(ref, pkg._ext.group_points(pts.to(cuda), idx.to(cuda)).cpu())
    gi = idx[:, :, 0].contiguous()
    ref = ref_ext.gather_points(pts.to(cuda), gi.to(cuda)).cpu()
    assert torch.equal(ref, oracle.ext.gather_points(pts, gi))
    assert torch.equal(ref, pkg._ext.gather_points(pts.to(cuda), gi.to(cuda)).cpu())
    i3 = torch.randint(0, 700, (2, 90, 3), generator=g, dtype=torch.int32)
    w3 = torch.rand(2, 90, 3, generator=g)
    ref = ref_ext.three_interpolate(pts.to(cuda), i3.to(cuda), w3.to(cuda)).cpu()
    assert torch.equal(ref, oracle.ext.three_interpolate(pts, i3, w3))
    assert torch.equal(ref, pkg._ext.three_interpolate(pts.to(cuda), i3.to(cuda), w3.to(cuda)).cpu())
    go = torch.randn(2, 19, 50, 8, generator=g)
    ref = ref_ext.group_points_grad(go.to(cuda), idx.to(cuda), 700).cpu()
    torch.testing.assert_close(ref, oracle.ext.group_points_grad(go, idx, 700), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ref, pkg._ext.group_points_grad(go.to(cuda), idx.to(cuda), 700).cpu(), rtol=1e-5, atol=1e-5)
