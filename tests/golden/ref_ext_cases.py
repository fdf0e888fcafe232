"""Seeded inputs of the reference-extension golden vectors (tests/golden/reference_ext.npz).

Shared by make_golden_ref_ext.py (run on the B200 box against the UNMODIFIED reference CUDA extension,
oracle/_ref/pointnet2_ref_ext.so) and by tests/test_oracle_vs_ref_ext_golden.py (replays the same inputs
through the CPU oracle without a GPU).  Only numpy's MT19937 and closed-form lattices: the inputs are
bit-identical on any machine, so the fixture stores OUTPUTS only."""
import numpy as np
import torch

from common import cloud


def _lattice(k):
    g = torch.stack(torch.meshgrid(*[torch.arange(float(k))] * 3, indexing="ij"), -1)
    return g.reshape(1, -1, 3).contiguous()


def fps_cases():
    """name -> (xyz (B,N,3), nsamples): ties (duplicates, lattice), the origin skip, n < 512 block sizes,
    ScanNet-sized clouds."""
    return {
        "uniform": (cloud(1, 2, 5000)[..., :3].contiguous(), 256),
        "dups": (cloud(2, 2, 4000, dup_frac=0.5)[..., :3].contiguous(), 512),
        "lattice": (_lattice(12), 700),
        "origin": (cloud(3, 1, 2048, origin_pts=40)[..., :3].contiguous(), 64),
        "n300": (cloud(6, 1, 300)[..., :3].contiguous(), 150),
        "n1000": (cloud(5, 1, 1000)[..., :3].contiguous(), 333),
        "n37": (cloud(7, 2, 37)[..., :3].contiguous(), 37),
        "n40000": (cloud(8, 1, 40000, extent=4.0, dup_frac=0.2, origin_pts=5)[..., :3].contiguous(), 2048),
    }


BALL_CFG = ((0.2, 64), (0.4, 32), (0.8, 16), (1.2, 16), (0.3, 16))   # the five (radius, nsample) of the path


def search_inputs():
    """xyz (2,6000,3) with duplicates, 256 centres taken from the cloud, 1024 `unknown` points."""
    x = cloud(90, 2, 6000, extent=1.0, dup_frac=0.1)[..., :3].contiguous()
    sel = np.random.RandomState(91).choice(6000, 256, replace=False)
    c = x[:, torch.from_numpy(sel)].contiguous()
    u = x[:, :1024].contiguous()
    return x, c, u


def index_inputs():
    rs = np.random.RandomState(5)
    pts = torch.from_numpy(rs.standard_normal((2, 19, 700)).astype(np.float32))
    idx = torch.from_numpy(rs.randint(0, 700, (2, 50, 8)).astype(np.int32))
    i3 = torch.from_numpy(rs.randint(0, 700, (2, 90, 3)).astype(np.int32))
    w3 = torch.from_numpy(rs.uniform(0, 1, (2, 90, 3)).astype(np.float32))
    go = torch.from_numpy(rs.standard_normal((2, 19, 50, 8)).astype(np.float32))
    return pts, idx, i3, w3, go


def run_all(ext, dev):
    """Every op of `ext` (the reference extension on `dev`, or the oracle's CPU stand-in) on the cases
    above -> dict of numpy arrays."""
    out = {}
    t = lambda a: a.to(dev)
    for name, (x, m) in fps_cases().items():
        out["fps_" + name] = ext.furthest_point_sampling(t(x), m).cpu().numpy()
    x, c, u = search_inputs()
    for r, ns in BALL_CFG:
        out["ball_%g_%d" % (r, ns)] = ext.ball_query(t(c), t(x), r, ns).cpu().numpy()
    d2, i = ext.three_nn(t(u), t(c))
    out["three_nn_dist2"], out["three_nn_idx"] = d2.cpu().numpy(), i.cpu().numpy()
    pts, idx, i3, w3, go = index_inputs()
    out["group_points"] = ext.group_points(t(pts), t(idx)).cpu().numpy()
    out["gather_points"] = ext.gather_points(t(pts), t(idx[:, :, 0].contiguous())).cpu().numpy()
    out["three_interpolate"] = ext.three_interpolate(t(pts), t(i3), t(w3)).cpu().numpy()
    out["group_points_grad"] = ext.group_points_grad(t(go), t(idx), 700).cpu().numpy()
    out["gather_points_grad"] = ext.gather_points_grad(t(go[..., 0].contiguous()), t(idx[:, :, 0].contiguous()), 700).cpu().numpy()
    return out
