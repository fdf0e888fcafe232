"""Drop-in for the reference's pybind11 module ``pointnet2._ext``.

Same nine names, argument order, dtypes, checks and allocation behaviour as
/root/reference/lib/pointnet2/_ext_src/src/bindings.cpp:11-24 and the wrappers in
src/{sampling,ball_query,group_points,interpolate}.cpp; the work is done by the sm_100a kernels of
lib3dvlp_b200.so through the C ABI (include/pointnet2_b200.h) on the current CUDA stream.
``sys.modules['pointnet2._ext'] = 3dvlp_b200._ext`` makes the reference's own
pointnet2_utils.py run on these kernels unchanged.
"""
import ctypes

import torch

from . import _lib as L

_f = ctypes.c_float


def furthest_point_sampling(points, nsamples):
    """sampling.cpp:70-91.  points (B,N,3) f32 -> (B,nsamples) i32."""
    L.check_contiguous(points, "points"); L.check_float(points, "points")
    L.require_cuda_path(points)
    b, n = points.size(0), points.size(1)
    out = torch.empty(b, nsamples, device=points.device, dtype=torch.int32)  # every index is written
    L.call("pn2_furthest_point_sampling", b, n, int(nsamples), L.ptr(points), None, L.ptr(out),
           L.stream_of(points), device=points.device)
    return out


def furthest_point_sampling_gather(points, nsamples, ordered=False):
    """FPS fused with the gather that follows it in every SA module: returns (inds, new_xyz).
    ordered=True: the caller believes `points` is the output of a previous FPS (SA2..SA4 of the
    backbone); the kernel side verifies that and short-cuts -- the result is the same either way."""
    L.check_contiguous(points, "points"); L.check_float(points, "points")
    L.require_cuda_path(points)
    b, n = points.size(0), points.size(1)
    out = torch.empty(b, nsamples, device=points.device, dtype=torch.int32)  # every index is written
    new_xyz = torch.empty(b, nsamples, 3, device=points.device, dtype=torch.float32)
    L.call("pn2_furthest_point_sampling_gather_ordered" if ordered else "pn2_furthest_point_sampling_gather",
           b, n, int(nsamples), L.ptr(points), L.ptr(out),
           L.ptr(new_xyz), L.stream_of(points), device=points.device)
    return out, new_xyz


def gather_points(points, idx):
    """sampling.cpp:20-43.  points (B,C,N) f32, idx (B,M) i32 -> (B,C,M)."""
    L.check_contiguous(points, "points"); L.check_contiguous(idx, "idx")
    L.check_float(points, "points"); L.check_int(idx, "idx")
    L.require_cuda_path(points); L.check_cuda(idx, "idx")
    b, c, n = points.shape
    m = idx.size(1)
    out = torch.empty(b, c, m, device=points.device, dtype=torch.float32)  # every element is written
    L.call("pn2_gather_points", b, c, n, m, L.ptr(points), L.ptr(idx), L.ptr(out), L.stream_of(points),
           device=points.device)
    return out


def gather_points_grad(grad_out, idx, n):
    """sampling.cpp:45-68.  grad_out (B,C,M), idx (B,M) -> (B,C,n) scatter-add."""
    L.check_contiguous(grad_out, "grad_out"); L.check_contiguous(idx, "idx")
    L.check_float(grad_out, "grad_out"); L.check_int(idx, "idx")
    L.require_cuda_path(grad_out); L.check_cuda(idx, "idx")
    b, c, m = grad_out.shape
    out = torch.zeros(b, c, n, device=grad_out.device, dtype=torch.float32)
    L.call("pn2_gather_points_grad", b, c, int(n), m, L.ptr(grad_out), L.ptr(idx), L.ptr(out),
           L.stream_of(grad_out), device=grad_out.device)
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """ball_query.cpp:13-37.  new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) i32."""
    L.check_contiguous(new_xyz, "new_xyz"); L.check_contiguous(xyz, "xyz")
    L.check_float(new_xyz, "new_xyz"); L.check_float(xyz, "xyz")
    L.require_cuda_path(new_xyz); L.check_cuda(xyz, "xyz")
    b, m = new_xyz.size(0), new_xyz.size(1)
    n = xyz.size(1)
    # the reference relies on torch::zeros for empty balls (ball_query.cpp:24-26); pn2_ball_query writes
    # those zeros itself (include/pointnet2_b200.h), so no fill kernel runs ahead of it
    idx = torch.empty(b, m, nsample, device=new_xyz.device, dtype=torch.int32)
    L.call("pn2_ball_query", b, n, m, _f(radius), int(nsample), L.ptr(new_xyz), L.ptr(xyz), L.ptr(idx),
           L.stream_of(new_xyz), device=new_xyz.device)
    return idx


def group_points(points, idx):
    """group_points.cpp:17-40.  points (B,C,N), idx (B,M,S) i32 -> (B,C,M,S)."""
    L.check_contiguous(points, "points"); L.check_contiguous(idx, "idx")
    L.check_float(points, "points"); L.check_int(idx, "idx")
    L.require_cuda_path(points); L.check_cuda(idx, "idx")
    b, c, n = points.shape
    m, s = idx.size(1), idx.size(2)
    # the reference zero-fills (group_points.cpp:27-29) and then overwrites every element; skipping the
    # fill saves a full pass over the largest tensor of the SA module
    out = torch.empty(b, c, m, s, device=points.device, dtype=torch.float32)
    L.call("pn2_group_points", b, c, n, m, s, L.ptr(points), L.ptr(idx), L.ptr(out), L.stream_of(points),
           device=points.device)
    return out


def group_points_grad(grad_out, idx, n):
    """group_points.cpp:42-65.  grad_out (B,C,M,S) -> (B,C,n) scatter-add."""
    L.check_contiguous(grad_out, "grad_out"); L.check_contiguous(idx, "idx")
    L.check_float(grad_out, "grad_out"); L.check_int(idx, "idx")
    L.require_cuda_path(grad_out); L.check_cuda(idx, "idx")
    b, c, m, s = grad_out.shape
    out = torch.zeros(b, c, n, device=grad_out.device, dtype=torch.float32)
    L.call("pn2_group_points_grad", b, c, int(n), m, s, L.ptr(grad_out), L.ptr(idx), L.ptr(out),
           L.stream_of(grad_out), device=grad_out.device)
    return out


def three_nn(unknowns, knows):
    """interpolate.cpp:19-45.  -> [dist2 (B,n,3) f32 SQUARED, idx (B,n,3) i32]."""
    L.check_contiguous(unknowns, "unknowns"); L.check_contiguous(knows, "knows")
    L.check_float(unknowns, "unknowns"); L.check_float(knows, "knows")
    L.require_cuda_path(unknowns); L.check_cuda(knows, "knows")
    b, n = unknowns.size(0), unknowns.size(1)
    m = knows.size(1)
    idx = torch.zeros(b, n, 3, device=unknowns.device, dtype=torch.int32)
    dist2 = torch.zeros(b, n, 3, device=unknowns.device, dtype=torch.float32)
    L.call("pn2_three_nn", b, n, m, L.ptr(unknowns), L.ptr(knows), L.ptr(dist2), L.ptr(idx),
           L.stream_of(unknowns), device=unknowns.device)
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """interpolate.cpp:47-75.  points (B,C,m), idx/weight (B,n,3) -> (B,C,n)."""
    L.check_contiguous(points, "points"); L.check_contiguous(idx, "idx"); L.check_contiguous(weight, "weight")
    L.check_float(points, "points"); L.check_int(idx, "idx"); L.check_float(weight, "weight")
    L.require_cuda_path(points); L.check_cuda(idx, "idx"); L.check_cuda(weight, "weight")
    b, c, m = points.shape
    n = idx.size(1)
    out = torch.empty(b, c, n, device=points.device, dtype=torch.float32)  # every element is written
    L.call("pn2_three_interpolate", b, c, m, n, L.ptr(points), L.ptr(idx), L.ptr(weight), L.ptr(out),
           L.stream_of(points), device=points.device)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """interpolate.cpp:77-104 as INTENDED (scatter-add); the reference binding launches the forward
    kernel here by mistake (interpolate.cpp:95)."""
    L.check_contiguous(grad_out, "grad_out"); L.check_contiguous(idx, "idx"); L.check_contiguous(weight, "weight")
    L.check_float(grad_out, "grad_out"); L.check_int(idx, "idx"); L.check_float(weight, "weight")
    L.require_cuda_path(grad_out); L.check_cuda(idx, "idx"); L.check_cuda(weight, "weight")
    b, c, n = grad_out.shape
    out = torch.zeros(b, c, m, device=grad_out.device, dtype=torch.float32)
    L.call("pn2_three_interpolate_grad", b, c, n, int(m), L.ptr(grad_out), L.ptr(idx), L.ptr(weight),
           L.ptr(out), L.stream_of(grad_out), device=grad_out.device)
    return out
