"""Box decode and axis-aligned IoU / DIoU, resident on the GPU.

Mirrors the pieces of /root/reference/utils/box_util.py the detector and the grounding ("OID") loss
use -- ``box3d_diou_batch_tensor`` (:488-528), ``rotz_batch_pytorch`` (:410-429), ``get_3d_box_batch``
(:361-384) -- and ``ProposalModule.decode_pred_box``
(models/proposal_module/proposal_module_fcos.py:94-144).  The reference evaluates the DIoU with ~45
element-wise torch kernels per call from Python loops over (scene, sentence) and decodes the box
corners through ``.cpu().numpy()`` every forward; here the IoU / DIoU of every (scene, sentence,
proposal) triple is ONE kernel launch (csrc/box_ops.cu, bit-identical fp32 op tree, analytic
backward) and the corners never leave the device.
"""
import ctypes
import math

import torch
from torch.autograd import Function

from . import _lib as L


def _f32c(t):
    return t.contiguous().float()


class _DiouPairs(Function):
    @staticmethod
    def forward(ctx, c1, s1, c2, s2):
        L.require_cuda_path(c1)
        b, k = c1.shape[0], c1.shape[1]
        l = c2.shape[1]
        c1, s1, c2, s2 = _f32c(c1), _f32c(s1), _f32c(c2), _f32c(s2)
        iou = torch.empty(b, l, k, device=c1.device, dtype=torch.float32)
        diou = torch.empty_like(iou)
        L.call("pn2_box3d_diou_pairs", b, l, k, L.ptr(c1), L.ptr(s1), L.ptr(c2), L.ptr(s2), L.ptr(iou), L.ptr(diou),
               L.stream_of(c1), device=c1.device)
        ctx.save_for_backward(c1, s1, c2, s2)
        return iou, diou

    @staticmethod
    def backward(ctx, g_iou, g_diou):
        c1, s1, c2, s2 = ctx.saved_tensors
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            raise RuntimeError("box3d_diou_pairs: gradients flow to the proposals only (targets are labels); "
                               "use box3d_diou_batch_tensor for a gradient w.r.t. both boxes")
        b, k = c1.shape[0], c1.shape[1]
        l = c2.shape[1]
        gc = torch.empty_like(c1)
        gs = torch.empty_like(s1)
        gi = _f32c(g_iou) if g_iou is not None else None
        gd = _f32c(g_diou) if g_diou is not None else None
        L.call("pn2_box3d_diou_pairs_grad", b, l, k, L.ptr(c1), L.ptr(s1), L.ptr(c2), L.ptr(s2), L.ptr(gi), L.ptr(gd),
               L.ptr(gc), L.ptr(gs), L.stream_of(c1), device=c1.device)
        return gc, gs, None, None


class _DiouElementwise(Function):
    @staticmethod
    def forward(ctx, c1, s1, c2, s2):
        L.require_cuda_path(c1)
        p = c1.shape[0]
        c1, s1, c2, s2 = _f32c(c1), _f32c(s1), _f32c(c2), _f32c(s2)
        iou = torch.empty(p, device=c1.device, dtype=torch.float32)
        diou = torch.empty_like(iou)
        L.call("pn2_box3d_diou_pairs", p, 1, 1, L.ptr(c1), L.ptr(s1), L.ptr(c2), L.ptr(s2), L.ptr(iou), L.ptr(diou),
               L.stream_of(c1), device=c1.device)
        ctx.save_for_backward(c1, s1, c2, s2)
        return iou, diou

    @staticmethod
    def backward(ctx, g_iou, g_diou):
        c1, s1, c2, s2 = ctx.saved_tensors
        need = ctx.needs_input_grad
        outs = [torch.empty_like(t) if n else None for t, n in zip((c1, s1, c2, s2), need)]
        gi = _f32c(g_iou) if g_iou is not None else None
        gd = _f32c(g_diou) if g_diou is not None else None
        L.call("pn2_box3d_diou_grad", ctypes.c_longlong(c1.shape[0]), L.ptr(c1), L.ptr(s1), L.ptr(c2), L.ptr(s2),
               L.ptr(gi), L.ptr(gd), L.ptr(outs[0]), L.ptr(outs[1]), L.ptr(outs[2]), L.ptr(outs[3]),
               L.stream_of(c1), device=c1.device)
        return tuple(outs)


def box3d_diou_batch_tensor(center1, size1, center2, size2):
    """box_util.py:488-528: element-wise IoU and clamped DIoU of P axis-aligned box pairs given as
    centre (P,3) + size (P,3).  Returns (iou (P,), diou (P,)), differentiable w.r.t. all four inputs."""
    return _DiouElementwise.apply(center1, size1, center2, size2)


def box3d_diou_pairs(pred_center, pred_size, gt_center, gt_size):
    """All proposals against all targets of their scene in one launch: pred_* (B,K,3), gt_* (B,L,3)
    -> (iou, diou) of shape (B,L,K), what loss_grounding.py:205-231 builds row by row with
    ``.repeat`` + box3d_diou_batch_tensor.  Differentiable w.r.t. the proposals."""
    return _DiouPairs.apply(pred_center, pred_size, gt_center, gt_size)


def rotz_batch_pytorch(t):
    """box_util.py:410-429 (transposed rotation about z, for x' A' = (A x)')."""
    c, s = torch.cos(t), torch.sin(t)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([c, -s, z, s, c, z, z, z, o], dim=-1).view(*t.shape, 3, 3)


def roty_batch_pytorch(t):
    """box_util.py:387-405."""
    c, s = torch.cos(t), torch.sin(t)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([c, z, s, z, o, z, -s, z, c], dim=-1).view(*t.shape, 3, 3)


_CORNER_SIGNS = ((1, 1, 1), (1, -1, 1), (-1, -1, 1), (-1, 1, 1), (1, 1, -1), (1, -1, -1), (-1, -1, -1), (-1, 1, -1))


_SIGNS_CACHE = {}


def _corner_signs(device, dtype):
    """The (8,3) sign table, uploaded once per device: a host->device copy of a Python list inside a CUDA
    graph capture would be illegal, and the detector's box decode is captured."""
    key = (str(device), dtype)
    t = _SIGNS_CACHE.get(key)
    if t is None:
        t = _SIGNS_CACHE[key] = torch.tensor(_CORNER_SIGNS, dtype=dtype).to(device)
    return t


def get_3d_box_batch(box_size, heading_angle, center):
    """box_util.py:361-384 on the device: box_size (...,3), heading_angle (...), center (...,3) ->
    corners (...,8,3).  Keeps the reference's quirk of rotating with ``roty_batch`` over corners laid
    out (l, w, h) (its own comment: "Actually we should rotz"); ScanNet headings are 0."""
    signs = _corner_signs(box_size.device, box_size.dtype)                                   # (8,3)
    corners = box_size.unsqueeze(-2) * 0.5 * signs                                           # (...,8,3)
    R = roty_batch_pytorch(heading_angle.to(box_size.dtype))                                 # (...,3,3)
    return torch.matmul(corners, R.transpose(-1, -2)) + center.unsqueeze(-2)


def decode_pred_box(data_dict, num_heading_bin):
    """proposal_module_fcos.py:94-144 without the host round trip: reads ``aggregated_vote_xyz``,
    ``heading_scores``, ``heading_residuals``, ``rois`` (distances to the six faces) and writes
    ``pred_heading``, ``pred_size``, ``pred_center``, ``pred_bbox_corner`` (the training-only random
    box masking of :161-178 stays with the caller)."""
    xyz = data_dict["aggregated_vote_xyz"]
    cls = torch.argmax(data_dict["heading_scores"], -1)
    res = torch.gather(data_dict["heading_residuals"], 2, cls.unsqueeze(-1))
    rois = data_dict["rois"].contiguous()
    pred_heading = cls.float() * (2.0 * math.pi / num_heading_bin) + res[..., 0]
    pred_size = rois[:, :, 0:3] + rois[:, :, 3:6]
    offset = (rois[:, :, 0:3] - rois[:, :, 3:6]) / 2
    R = rotz_batch_pytorch(pred_heading.float())
    offset = torch.matmul(offset.unsqueeze(-2), R).squeeze(-2)
    pred_center = xyz - offset
    data_dict["pred_heading"] = pred_heading
    data_dict["pred_size"] = pred_size
    data_dict["pred_center"] = pred_center
    data_dict["pred_bbox_corner"] = get_3d_box_batch(pred_size.detach(), pred_heading.detach(), pred_center.detach())
    return data_dict
