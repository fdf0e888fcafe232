"""CPU oracle for the PointNet++ op suite -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The product package (3dvlp_b200/) never does, and has no CPU fallback.

Two layers:

* ``ext`` -- an object with the nine function names of the reference's ``pointnet2._ext`` pybind
  module (/root/reference/lib/pointnet2/_ext_src/src/bindings.cpp:11-24), same argument order,
  dtypes and allocation behaviour (zero-filled outputs), running on CPU tensors through
  oracle/pointnet2_oracle.c.  Injecting it as ``pointnet2_utils._ext`` makes the reference's own
  Python layers run on CPU (that is how tests/golden/make_golden.py produces the fixtures).
* functional restatements of the Python layers above it: ``query_and_group``
  (pointnet2_utils.py:313-372), ``sa_module`` (pointnet2_modules.py:210-272), ``fp_module``
  (:371-416), ``shared_mlp`` (pytorch_utils.py:11-36), ``backbone`` (backbone_module.py:76-135),
  ``nn_distance`` (utils/nn_distance.py:32-59), in plain fp32 torch on CPU.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    """Compile oracle/pointnet2_oracle.c -> oracle/liboracle.so (gcc; see oracle/Makefile)."""
    src = os.path.join(_HERE, "pointnet2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_opt_n_threads.restype = ctypes.c_int
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _chk(t, dtype, name):
    if not t.is_contiguous():
        raise RuntimeError(name + " must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(name + (" must be a float tensor" if dtype == torch.float32 else " must be an int tensor"))


N_THREADS = max(1, os.cpu_count() or 1)
_pool = None


def _run_chunks(jobs, threads):
    """Run ctypes calls (which release the GIL) on a thread pool; `jobs` = list of thunks."""
    global _pool
    if threads <= 1 or len(jobs) <= 1:
        for j in jobs:
            j()
        return
    if _pool is None or _pool._max_workers < threads:
        _pool = ThreadPoolExecutor(max_workers=threads)
    list(_pool.map(lambda f: f(), jobs))


class _Ext:
    """CPU stand-in for the reference's ``pointnet2._ext`` (bindings.cpp:11-24)."""

    threads = 1  # set to N_THREADS for the cpu-baseline timing legs

    def furthest_point_sampling(self, points, nsamples):
        _chk(points, torch.float32, "points")
        b, n, _ = points.shape
        out = torch.zeros(b, nsamples, dtype=torch.int32)
        tmp = torch.full((b, n), 1e10, dtype=torch.float32)
        L = lib()
        jobs = [
            (lambda i=i: L.oracle_furthest_point_sampling(
                1, n, nsamples, _p(points[i]), _p(tmp[i]), _p(out[i])))
            for i in range(b)
        ]
        _run_chunks(jobs, self.threads)
        return out

    def gather_points(self, points, idx):
        _chk(points, torch.float32, "points"); _chk(idx, torch.int32, "idx")
        b, c, n = points.shape
        m = idx.shape[1]
        out = torch.zeros(b, c, m, dtype=torch.float32)
        lib().oracle_gather_points(b, c, n, m, _p(points), _p(idx), _p(out))
        return out

    def gather_points_grad(self, grad_out, idx, n):
        _chk(grad_out, torch.float32, "grad_out"); _chk(idx, torch.int32, "idx")
        b, c, m = grad_out.shape
        out = torch.zeros(b, c, n, dtype=torch.float32)
        lib().oracle_gather_points_grad(b, c, n, m, _p(grad_out), _p(idx), _p(out))
        return out

    def ball_query(self, new_xyz, xyz, radius, nsample):
        _chk(new_xyz, torch.float32, "new_xyz"); _chk(xyz, torch.float32, "xyz")
        b, m, _ = new_xyz.shape
        n = xyz.shape[1]
        idx = torch.zeros(b, m, nsample, dtype=torch.int32)
        L = lib()
        r = ctypes.c_float(radius)
        chunk = max(1, -(-m // max(1, self.threads // max(1, b) + 1)))
        jobs = []
        for i in range(b):
            for j0 in range(0, m, chunk):
                j1 = min(m, j0 + chunk)
                jobs.append(lambda i=i, j0=j0, j1=j1: L.oracle_ball_query(
                    1, n, j1 - j0, r, nsample, _p(new_xyz[i, j0:j1]), _p(xyz[i]), _p(idx[i, j0:j1])))
        _run_chunks(jobs, self.threads)
        return idx

    def group_points(self, points, idx):
        _chk(points, torch.float32, "points"); _chk(idx, torch.int32, "idx")
        b, c, n = points.shape
        _, npoints, nsample = idx.shape
        out = torch.zeros(b, c, npoints, nsample, dtype=torch.float32)
        lib().oracle_group_points(b, c, n, npoints, nsample, _p(points), _p(idx), _p(out))
        return out

    def group_points_grad(self, grad_out, idx, n):
        _chk(grad_out, torch.float32, "grad_out"); _chk(idx, torch.int32, "idx")
        b, c, npoints, nsample = grad_out.shape
        out = torch.zeros(b, c, n, dtype=torch.float32)
        lib().oracle_group_points_grad(b, c, n, npoints, nsample, _p(grad_out), _p(idx), _p(out))
        return out

    def three_nn(self, unknowns, knows):
        _chk(unknowns, torch.float32, "unknowns"); _chk(knows, torch.float32, "knows")
        b, n, _ = unknowns.shape
        m = knows.shape[1]
        dist2 = torch.zeros(b, n, 3, dtype=torch.float32)
        idx = torch.zeros(b, n, 3, dtype=torch.int32)
        lib().oracle_three_nn(b, n, m, _p(unknowns), _p(knows), _p(dist2), _p(idx))
        return dist2, idx

    def three_interpolate(self, points, idx, weight):
        _chk(points, torch.float32, "points"); _chk(idx, torch.int32, "idx")
        _chk(weight, torch.float32, "weight")
        b, c, m = points.shape
        n = idx.shape[1]
        out = torch.zeros(b, c, n, dtype=torch.float32)
        lib().oracle_three_interpolate(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out))
        return out

    def three_interpolate_grad(self, grad_out, idx, weight, m):
        _chk(grad_out, torch.float32, "grad_out"); _chk(idx, torch.int32, "idx")
        _chk(weight, torch.float32, "weight")
        b, c, n = grad_out.shape
        out = torch.zeros(b, c, m, dtype=torch.float32)
        lib().oracle_three_interpolate_grad(b, c, n, m, _p(grad_out), _p(idx), _p(weight), _p(out))
        return out


ext = _Ext()


def opt_n_threads(work_size):
    return int(lib().oracle_opt_n_threads(int(work_size)))


def nn_distance(pc1, pc2, l1smooth=False, delta=1.0, l1=False):
    """utils/nn_distance.py:32-59 through the C loop (fp32, first-index ties, int64 indices)."""
    pc1 = pc1.contiguous().float(); pc2 = pc2.contiguous().float()
    b, n, c = pc1.shape
    m = pc2.shape[1]
    d1 = torch.empty(b, n); i1 = torch.empty(b, n, dtype=torch.int64)
    d2 = torch.empty(b, m); i2 = torch.empty(b, m, dtype=torch.int64)
    mode = 2 if l1smooth else (1 if l1 else 0)
    lib().oracle_nn_distance(b, n, m, c, _p(pc1), _p(pc2), mode, ctypes.c_float(delta),
                             _p(d1), _p(i1), _p(d2), _p(i2))
    return d1, i1, d2, i2


def nn_distance_torch(pc1, pc2, l1smooth=False, delta=1.0, l1=False):
    """Literal torch restatement of utils/nn_distance.py:45-59 (differentiable; CPU or GPU)."""
    n, m = pc1.shape[1], pc2.shape[1]
    diff = pc1.unsqueeze(2).repeat(1, 1, m, 1) - pc2.unsqueeze(1).repeat(1, n, 1, 1)
    if l1smooth:
        a = diff.abs()
        q = torch.clamp(a, max=delta)
        dist = (0.5 * q ** 2 + delta * (a - q)).sum(-1)
    elif l1:
        dist = diff.abs().sum(-1)
    else:
        dist = (diff ** 2).sum(-1)
    d1, i1 = torch.min(dist, dim=2)
    d2, i2 = torch.min(dist, dim=1)
    return d1, i1, d2, i2


# ----------------------------------------------------------------------------------------------
# functional restatement of the Python layers (fp32 torch on CPU)
# ----------------------------------------------------------------------------------------------

def query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz=True, normalize_xyz=False):
    """pointnet2_utils.py:313-372 (sample_uniformly off).  Returns (new_features, grouped_xyz, idx)."""
    idx = ext.ball_query(new_xyz.contiguous(), xyz.contiguous(), radius, nsample)
    grouped_xyz = ext.group_points(xyz.transpose(1, 2).contiguous(), idx)
    grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        grouped_xyz = grouped_xyz / radius
    if features is not None:
        gf = ext.group_points(features.contiguous(), idx)
        new_features = torch.cat([grouped_xyz, gf], dim=1) if use_xyz else gf
    else:
        new_features = grouped_xyz
    return new_features, grouped_xyz, idx


def shared_mlp(x, sd, prefix, n_layers, training=False, bn=True, eps=1e-5):
    """pytorch_utils.py:11-36: per layer Conv2d 1x1 (no bias when bn) -> BatchNorm2d -> ReLU.

    `sd` is a state_dict with the reference's key names (``{prefix}layer{i}.conv.weight`` ...).
    Eval mode uses the running statistics; training mode uses biased batch statistics.
    """
    for i in range(n_layers):
        w = sd[f"{prefix}layer{i}.conv.weight"]
        bias = sd.get(f"{prefix}layer{i}.conv.bias")
        x = F.conv2d(x, w, bias)
        if bn:
            p = f"{prefix}layer{i}.bn.bn."
            x = F.batch_norm(x, None if training else sd[p + "running_mean"],
                             None if training else sd[p + "running_var"],
                             sd[p + "weight"], sd[p + "bias"], training=training, eps=eps)
        x = F.relu(x)
    return x


def sa_module(xyz, features, sd, prefix, npoint, radius, nsample, n_layers=3, use_xyz=True,
              normalize_xyz=False, pooling="max", sigma=None, inds=None, training=False):
    """PointnetSAModuleVotes.forward, pointnet2_modules.py:210-272."""
    xyz = xyz.contiguous()
    xyz_flipped = xyz.transpose(1, 2).contiguous()
    if inds is None:
        inds = ext.furthest_point_sampling(xyz, npoint)
    new_xyz = ext.gather_points(xyz_flipped, inds).transpose(1, 2).contiguous()
    grouped, grouped_xyz, idx = query_and_group(xyz, new_xyz, features, radius, nsample,
                                                use_xyz=use_xyz, normalize_xyz=normalize_xyz)
    y = shared_mlp(grouped, sd, prefix + "mlp_module.", n_layers, training=training)
    if pooling == "max":
        y = F.max_pool2d(y, kernel_size=[1, y.size(3)])
    elif pooling == "avg":
        y = F.avg_pool2d(y, kernel_size=[1, y.size(3)])
    elif pooling == "rbf":
        if sigma is None:
            sigma = radius / 2
        rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (sigma ** 2) / 2)
        y = torch.sum(y * rbf.unsqueeze(1), -1, keepdim=True) / float(nsample)
    return new_xyz, y.squeeze(-1), inds


def fp_module(unknown, known, unknow_feats, known_feats, sd, prefix, n_layers=2, training=False):
    """PointnetFPModule.forward, pointnet2_modules.py:371-416."""
    if known is not None:
        dist2, idx = ext.three_nn(unknown.contiguous(), known.contiguous())
        dist = torch.sqrt(dist2)
        dist_recip = 1.0 / (dist + 1e-8)
        norm = torch.sum(dist_recip, dim=2, keepdim=True)
        weight = dist_recip / norm
        interp = ext.three_interpolate(known_feats.contiguous(), idx, weight.contiguous())
    else:
        interp = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
    x = torch.cat([interp, unknow_feats], dim=1) if unknow_feats is not None else interp
    x = shared_mlp(x.unsqueeze(-1), sd, prefix + "mlp.", n_layers, training=training)
    return x.squeeze(-1)


def _bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def fold_bn(sd, prefix, i, eps=1e-5):
    """Conv + eval-mode BatchNorm of layer i folded into (W' (out,in), b' (out)), fp32."""
    w = sd[f"{prefix}layer{i}.conv.weight"].float()
    w = w.reshape(w.shape[0], -1)
    b = sd.get(f"{prefix}layer{i}.conv.bias")
    b = b.float() if b is not None else torch.zeros(w.shape[0])
    p = f"{prefix}layer{i}.bn.bn."
    if p + "weight" in sd:
        scale = sd[p + "weight"].float() / torch.sqrt(sd[p + "running_var"].float() + eps)
        w = w * scale[:, None]
        b = (b - sd[p + "running_mean"].float()) * scale + sd[p + "bias"].float()
    return w, b


def sa_module_mixed(xyz, features, sd, prefix, npoint, radius, nsample, normalize_xyz=True, inds=None):
    """The fused SA path's numerics contract restated in torch (eval mode, 3-layer MLP, max pool):
    bf16-rounded operands for every tensor-core contraction, exact products, fp32 accumulation;
    geometric half of layer 1 and <= 4 raw feature channels in fp32.  See 3dvlp_b200/fused.py."""
    xyz = xyz.contiguous()
    if inds is None:
        inds = ext.furthest_point_sampling(xyz, npoint)
    new_xyz = ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    idx = ext.ball_query(new_xyz, xyz, radius, nsample)
    gx = ext.group_points(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        gx = gx / radius
    w1, b1 = fold_bn(sd, prefix + "mlp_module.", 0)
    w2, b2 = fold_bn(sd, prefix + "mlp_module.", 1)
    w3, b3 = fold_bn(sd, prefix + "mlp_module.", 2)
    y = torch.einsum("oc,bcms->boms", w1[:, :3].double(), gx.double()) + b1.double()[None, :, None, None]
    c_in = w1.shape[1] - 3
    if c_in:
        if c_in <= 4:
            gf = ext.group_points(features.contiguous(), idx)
            y = y + torch.einsum("oc,bcms->boms", w1[:, 3:].double(), gf.double())
        else:
            g = torch.einsum("oc,bcn->bon", _bf16r(w1[:, 3:]).double(), _bf16r(features).double()).float()
            y = y + ext.group_points(g.contiguous(), idx).double()
    x1 = _bf16r(torch.relu(y).float())
    y2 = torch.einsum("oc,bcms->boms", _bf16r(w2).double(), x1.double()) + b2.double()[None, :, None, None]
    x2 = _bf16r(torch.relu(y2).float())
    y3 = torch.einsum("oc,bcms->boms", _bf16r(w3).double(), x2.double())
    out = torch.relu(y3.max(dim=3).values + b3.double()[None, :, None]).float()
    return new_xyz, out, inds


def fp_module_mixed(unknown, known, unknow_feats, known_feats, sd, prefix, n_layers=2):
    """Numerics contract of the fused FP path: fp32 3-NN / weights / interpolation, then every MLP
    layer as a bf16 x bf16 -> fp32 contraction with folded BatchNorm (hidden activations bf16)."""
    dist2, idx = ext.three_nn(unknown.contiguous(), known.contiguous())
    r = 1.0 / (torch.sqrt(dist2) + 1e-8)
    w = r / torch.sum(r, dim=2, keepdim=True)
    interp = ext.three_interpolate(known_feats.contiguous(), idx, w.contiguous())
    x = torch.cat([interp, unknow_feats], dim=1) if unknow_feats is not None else interp
    for i in range(n_layers):
        wgt, b = fold_bn(sd, prefix + "mlp.", i)
        x = torch.relu(torch.einsum("oc,bcn->bon", _bf16r(wgt).double(), _bf16r(x).double())
                       + b.double()[None, :, None]).float()
    return x


BACKBONE_SA = (  # backbone_module.py:29-63: (name, npoint, radius, nsample)
    ("sa1", 2048, 0.2, 64), ("sa2", 1024, 0.4, 32), ("sa3", 512, 0.8, 16), ("sa4", 256, 1.2, 16))


def backbone(pointcloud, sd, sa_cfg=BACKBONE_SA, training=False):
    """Pointnet2Backbone.forward, backbone_module.py:76-135, on a reference-keyed state_dict."""
    out = {}
    xyz = pointcloud[..., :3].contiguous()
    features = pointcloud[..., 3:].transpose(1, 2).contiguous() if pointcloud.size(-1) > 3 else None
    for name, npoint, radius, nsample in sa_cfg:
        xyz, features, inds = sa_module(xyz, features, sd, name + ".", npoint, radius, nsample,
                                        normalize_xyz=True, training=training)
        out[name + "_xyz"] = xyz
        out[name + "_features"] = features
        if name in ("sa1", "sa2"):
            out[name + "_inds"] = inds
    f = fp_module(out["sa3_xyz"], out["sa4_xyz"], out["sa3_features"], out["sa4_features"], sd,
                  "fp1.", training=training)
    f = fp_module(out["sa2_xyz"], out["sa3_xyz"], out["sa2_features"], f, sd, "fp2.",
                  training=training)
    out["fp2_features"] = f
    out["fp2_xyz"] = out["sa2_xyz"]
    num_seed = out["fp2_xyz"].shape[1]
    out["fp2_inds"] = out["sa1_inds"][:, 0:num_seed]
    return out


def furthest_point_sampling_keyed(points, nsamples):
    """FPS restated through the index-only tie-break key of SURVEY.md appendix A-1
    (``oracle_furthest_point_sampling_keyed`` in pointnet2_oracle.c): next sample =
    argmax_k of the packed (bits(temp[k]) << 32 | ~key32(k)).  tests/test_oracle.py checks it
    against the literal thread/tree simulation; the CUDA kernel implements this formulation."""
    _chk(points, torch.float32, "points")
    b, n, _ = points.shape
    out = torch.zeros(b, nsamples, dtype=torch.int32)
    tmp = torch.full((b, n), 1e10, dtype=torch.float32)
    lib().oracle_furthest_point_sampling_keyed(b, n, nsamples, _p(points), _p(tmp), _p(out))
    return out


def _terms(t, n):
    """v -> its first n bf16 terms (h = bf16(v), m = bf16(v - h), l = bf16(v - h - m)) as float64 tensors."""
    out, rest = [], t.float()
    for _ in range(n):
        h = _bf16r(rest)
        out.append(h.double())
        rest = rest - h
    return out


_PRODUCTS = {1: ((0, 0),), 2: ((0, 0), (1, 0), (0, 1)), 3: ((0, 0), (1, 0), (0, 1), (2, 0), (0, 2), (1, 1))}


def _contract(eq, a, b, terms):
    """einsum(eq, a, b) under the operand contract of 3dvlp_b200/mlp_train.py: both operands as `terms` bf16
    terms, the leading term products summed exactly."""
    at, bt = _terms(a, terms), _terms(b, terms)
    return sum(torch.einsum(eq, at[i], bt[j]) for i, j in _PRODUCTS[terms])


def shared_mlp_pool_mixed(x, weights, gammas, betas, running=None, training=True, eps=1e-5, momentum=0.1, dout=None,
                          terms=3):
    """Numerics contract of the training-mode SharedMLP + max-pool path (3dvlp_b200/mlp_train.py) restated in
    float64 torch with the same operand roundings: every operand of a contraction (input, weights, hidden
    activations, dY) as `terms` bf16 terms with the leading products (3: "bf16x6", 2: "bf16x3", 1: "bf16");
    everything else (BatchNorm statistics, ReLU, pooling, their gradients) unrounded.
    x (B, C0, M, S); weights[l] (Co, Ci); running = list of (running_mean, running_var) updated in place.
    Returns (out (B, C_last, M), grads) with grads = None or (dx, [dW], [dgamma], [dbeta]) for `dout`."""
    X = x.double()
    n = len(weights)
    saved = []
    P = x.shape[0] * x.shape[2] * x.shape[3]
    A = None
    for l in range(n):
        W = weights[l].double().reshape(weights[l].shape[0], -1)
        Y = _contract("oc,bcms->boms", W, X, terms).float().double()     # Y is kept in fp32 on the device
        if training:
            mean = Y.mean(dim=(0, 2, 3))
            var = Y.var(dim=(0, 2, 3), unbiased=False)
            if running is not None:
                rm, rv = running[l]
                rm.mul_(1 - momentum).add_(momentum * mean.float())
                rv.mul_(1 - momentum).add_(momentum * (var * P / max(P - 1, 1)).float())
        else:
            mean, var = running[l][0].double(), running[l][1].double()
        invstd = 1.0 / torch.sqrt(var + eps)
        xhat = (Y - mean[None, :, None, None]) * invstd[None, :, None, None]
        A = torch.relu(xhat * gammas[l].double()[None, :, None, None] + betas[l].double()[None, :, None, None])
        saved.append((X, W, xhat, invstd, A))
        X = A.float().double()
    out, arg = A.max(dim=3)   # first maximal slot, like F.max_pool2d
    if dout is None:
        return out.float(), None
    dA = torch.zeros_like(A).scatter_(3, arg.unsqueeze(3), dout.double().unsqueeze(3))
    dWs, dgs, dbs = [None] * n, [None] * n, [None] * n
    for l in range(n - 1, -1, -1):
        Xl, W, xhat, invstd, A = saved[l]
        dYhat = dA * (A > 0)
        s1 = dYhat.sum(dim=(0, 2, 3))
        s2 = (dYhat * xhat).sum(dim=(0, 2, 3))
        dgs[l], dbs[l] = s2.float(), s1.float()
        g = (gammas[l].double() * invstd)[None, :, None, None]
        if training:
            dY = g * (dYhat - (s1[None, :, None, None] + xhat * s2[None, :, None, None]) / P)
        else:
            dY = g * dYhat
        dY = dY.float().double()
        dWs[l] = _contract("boms,bcms->oc", dY, Xl, terms).float()
        dA = _contract("oc,boms->bcms", W, dY, terms)
    return out.float(), (dA.float(), dWs, dgs, dbs)
