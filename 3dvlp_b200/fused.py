"""Fused eval-mode execution of the SA / FP modules on the sm_100a kernels.

``sa_forward`` replaces the body of PointnetSAModuleVotes.forward (reference
pointnet2_modules.py:233-270) by four launches -- FPS+gather, ball query, (for wide inputs) one
tensor-core GEMM over the input points, and the fused grouping + shared-MLP + max-pool kernel --
with BatchNorm folded into the convolution weights (eval mode: running statistics, eps as in the
module).  Returning None tells the caller to use its generic op-by-op path (training, odd shapes).

Numerics of the fused path (the contract the tests assert, tests/test_fused_gpu.py):
every tensor-core operand (activations, folded weights) is rounded to bf16 (RNE); products are
exact and accumulate in fp32; the geometric half of layer 1 (3 FMAs per channel on the recentred,
radius-divided coordinates) and up to 4 raw feature channels are evaluated in fp32.
"""
import ctypes

import torch
import torch.nn as nn

from . import _ext
from . import _lib as L
from . import tensorcore as tcm

_NS_OK = (16, 32, 64, 128)


def _fold_block(block):
    """Conv (1x1, optional bias) + optional BatchNorm(eval) -> (W' (out,in), b' (out)) in fp32."""
    conv = block.conv
    w = conv.weight.detach().float().reshape(conv.weight.shape[0], -1)
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    if hasattr(block, "bn"):
        bn = block.bn.bn
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * scale[:, None]
        b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return w.contiguous(), b.contiguous()


def _plain_relu_mlp(mlp, n_layers):
    """True if `mlp` is a SharedMLP of exactly n_layers post-activation conv(+bn)+ReLU blocks."""
    blocks = list(mlp.children())
    if len(blocks) != n_layers:
        return False
    for blk in blocks:
        names = [n for n, _ in blk.named_children()]
        if names not in (["conv", "bn", "activation"], ["conv", "activation"]):
            return False
        if not isinstance(blk.activation, nn.ReLU):
            return False
        if tuple(blk.conv.kernel_size) != (1,) * len(blk.conv.kernel_size):
            return False
    return True


def _version_key(module):
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


MAX_NPOINT = 4096   # limits of pn2_sa_group_mlp's row plan (include/pointnet2_b200.h)
MAX_SCENES = 64


class _SaPlan:
    def __init__(self, module):
        blocks = list(module.mlp_module.children())
        w1, self.b1 = _fold_block(blocks[0])
        w2, self.b2 = _fold_block(blocks[1])
        w3, self.b3 = _fold_block(blocks[2])
        self.c1, self.c2, self.c3 = w1.shape[0], w2.shape[0], w3.shape[0]
        self.c_in = w1.shape[1] - 3
        self.direct = self.c_in <= 4
        if self.direct:
            self.w1 = w1.contiguous()
            self.w1f = None
        else:
            self.w1 = w1[:, :3].contiguous()
            self.w1f = tcm.PackedWeight(w1[:, 3:].contiguous())
        self.w2p = tcm.PackedWeight(w2)
        self.w3p = tcm.PackedWeight(w3)
        self.key = _version_key(module)

    @staticmethod
    def supported(module):
        if not (module.use_xyz and module.pooling == "max" and module.nsample in _NS_OK):
            return False
        if module.npoint is None or module.npoint > MAX_NPOINT:  # row plan of pn2_sa_group_mlp
            return False
        if not _plain_relu_mlp(module.mlp_module, 3):
            return False
        blocks = list(module.mlp_module.children())
        c1, c2, c3 = (b.conv.out_channels for b in blocks)
        if blocks[0].conv.in_channels < 3:
            return False
        return c1 in (64, 128) and c2 in (64, 128) and c3 in (128, 256)


def _sa_plan(module):
    """The cached folded/packed parameters of an SA module, or None if it cannot run fused."""
    if module._fused is False:
        return None
    if module._fused is None or module._fused.key != _version_key(module):
        if not _SaPlan.supported(module):
            module._fused = False
            return None
        module._fused = _SaPlan(module)
    return module._fused


def sa_sample(module, xyz, inds=None, ordered=False):
    """FPS (+ the xyz gather) of an SA layer -- geometry only, independent of the features."""
    if inds is None:
        return _ext.furthest_point_sampling_gather(xyz, module.npoint, ordered=ordered)
    new_xyz = torch.gather(xyz, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    return inds, new_xyz


def sa_features(module, plan, xyz, features, new_xyz, rows=None, idx=None):
    """Ball query + (wide inputs: per-point GEMM) + the fused grouping/MLP/max-pool kernel.
    `rows`: the (B*N, kpad) bf16 point-major operand of the per-point GEMM when the caller already
    has it (then `features` is not read)."""
    b, n = xyz.shape[0], xyz.shape[1]
    dev = xyz.device
    c_feat = 0 if features is None else features.shape[1]
    if idx is None:
        idx = _ext.ball_query(new_xyz, xyz, module.radius, module.nsample)
    g = None
    feat = None
    if plan.direct:
        if c_feat:
            feat = features.contiguous().float()
    else:
        if rows is None:
            rows = cached_rows(features)
        if rows is None or rows.shape != (b * n, plan.w1f.kpad) or rows.device != dev:
            rows = tcm.pack_rows(features.contiguous().float(), plan.w1f.kpad)
        g = tcm.linear(rows, plan.w1f, None, relu=False, out="f32")
    out_cm = torch.empty(b, plan.c3, module.npoint, device=dev, dtype=torch.float32)
    out_pm = torch.empty(b * module.npoint, plan.c3, device=dev, dtype=torch.bfloat16)
    radius_div = float(module.radius) if module.normalize_xyz else 1.0
    L.call("pn2_sa_group_mlp", b, n, module.npoint, module.nsample, ctypes.c_float(radius_div), L.ptr(xyz),
           L.ptr(new_xyz), L.ptr(idx), plan.c_in if plan.direct else 0, L.ptr(feat), L.ptr(g), plan.c1, plan.c2,
           plan.c3, L.ptr(plan.w1), L.ptr(plan.b1), L.ptr(plan.w2p.data), L.ptr(plan.b2), L.ptr(plan.w3p.data),
           L.ptr(plan.b3), L.ptr(out_cm), L.ptr(out_pm), L.stream_of(xyz), device=dev)
    attach_rows(out_cm, out_pm)  # bf16 point-major twin: the next SA layer's GEMM operand, no repack
    return out_cm


def attach_rows(features, rows):
    """Remember the bf16 point-major twin of `features` together with the tensor version it was made
    from: an in-place edit of the fp32 tensor (mul_, masking, copy_) invalidates the twin."""
    features._pn2_rows = (features._version, rows)


def cached_rows(features):
    """The twin attached by attach_rows, or None when there is none or `features` has changed since."""
    rec = getattr(features, "_pn2_rows", None)
    if rec is None or rec[0] != features._version:
        return None
    return rec[1]


def sa_forward(module, xyz, features, inds):
    plan = _sa_plan(module)
    if plan is None:
        return None
    c_feat = 0 if features is None else features.shape[1]
    if c_feat != plan.c_in or xyz.dtype != torch.float32 or xyz.shape[0] > MAX_SCENES:
        return None
    xyz = xyz.contiguous()
    inds, new_xyz = sa_sample(module, xyz, inds)
    return new_xyz, sa_features(module, plan, xyz, features, new_xyz), inds


_side_streams = {}
_aux_streams = {}


def backbone_forward(bb, data_dict):
    """Whole-backbone fast path (eval, no grad).  The sampling chain FPS1 -> FPS2 -> FPS3 -> FPS4 only
    needs coordinates, never features, and it is the latency-bound part of the step (each FPS is a
    serial chain that occupies a few SMs).  It therefore runs on a SIDE stream, and the feature path
    (ball query, grouped MLP, FP layers) of layer i runs on the caller's stream as soon as FPS i has
    finished -- underneath FPS i+1.  Same kernels, same results, shorter critical path.
    Returns None when any layer cannot run fused (the caller then goes module by module)."""
    import os
    if os.environ.get("PN2_TWO_STREAM", "1") == "0":
        return None
    sas = (bb.sa1, bb.sa2, bb.sa3, bb.sa4)
    plans = []
    for sa in sas:
        if not sa._can_fuse(data_dict["point_clouds"], None):
            return None
        plan = _sa_plan(sa)
        if plan is None:
            return None
        plans.append(plan)
    pc = data_dict["point_clouds"]
    if pc.dtype != torch.float32 or pc.size(-1) - 3 != plans[0].c_in or pc.shape[0] > MAX_SCENES:
        return None
    rows1 = None
    if plans[0].direct or plans[0].c_in == 0:
        xyz, features = bb._break_up_pc(pc)
    else:
        # wide inputs (multiview features): the per-point GEMM of SA1 reads point-major rows, which is
        # how the cloud arrives -- convert in place of _break_up_pc's transpose to (B, C, N) and back
        pc = pc.contiguous()
        xyz, features = pc[..., :3].contiguous(), None
        rows1 = tcm.pack_point_rows(pc, 3, plans[0].c_in, plans[0].w1f.kpad)
    dev = xyz.device
    main = torch.cuda.current_stream(dev)
    # the helper streams take the caller's stream priority: kernels captured into a CUDA graph keep the
    # priority of the stream they were captured on, and a serving loop uses it to stagger its batches
    prio = main.priority
    side = _side_streams.get((dev, prio))
    if side is None:
        side = _side_streams[(dev, prio)] = torch.cuda.Stream(device=dev, priority=prio)
    aux = _aux_streams.get((dev, prio))
    if aux is None:
        aux = _aux_streams[(dev, prio)] = torch.cuda.Stream(device=dev, priority=prio)
    capturing = torch.cuda.is_current_stream_capturing()
    side.wait_stream(main)
    aux.wait_stream(main)
    geo = []
    with torch.cuda.stream(side):
        x = xyz
        for li, sa in enumerate(sas):
            # from the second layer on the input is the previous layer's FPS output: almost surely in
            # furthest-point order already, which the kernel side verifies instead of recomputing
            inds, new_xyz = sa_sample(sa, x, ordered=li > 0)
            ev = torch.cuda.Event()
            ev.record(side)
            geo.append((inds, new_xyz, ev))
            x = new_xyz
        # the FP layers' neighbour search needs coordinates only: it runs here, under the SA feature kernels
        nn1 = fp_neighbours(geo[2][1], geo[3][1])   # fp1: unknown = sa3_xyz, known = sa4_xyz
        nn2 = fp_neighbours(geo[1][1], geo[2][1])   # fp2: unknown = sa2_xyz, known = sa3_xyz
        nn_ev = torch.cuda.Event()
        nn_ev.record(side)
    # ball queries need coordinates only too, but must not delay the sampling chain: a third stream runs
    # query i as soon as FPS i is done, next to FPS i+1 and to the feature kernels of layer i-1
    balls = []
    with torch.cuda.stream(aux):
        x = xyz
        for sa, (inds, new_xyz, ev) in zip(sas, geo):
            aux.wait_event(ev)
            if not capturing:
                new_xyz.record_stream(aux)
            idx = _ext.ball_query(new_xyz, x, sa.radius, sa.nsample)
            bev = torch.cuda.Event()
            bev.record(aux)
            balls.append((idx, bev))
            x = new_xyz
    x, f = xyz, features
    names = ("sa1", "sa2", "sa3", "sa4")
    for name, sa, plan, (inds, new_xyz, ev), (idx, bev) in zip(names, sas, plans, geo, balls):
        main.wait_event(ev)
        main.wait_event(bev)
        if not capturing:  # graph-private pool memory is never handed to another stream
            inds.record_stream(main)
            new_xyz.record_stream(main)
            idx.record_stream(main)
        r = rows1 if name == "sa1" else None
        c_feat = plan.c_in if r is not None else (0 if f is None else f.shape[1])
        if c_feat != plan.c_in:
            main.wait_stream(side)
            main.wait_stream(aux)
            return None
        f = sa_features(sa, plan, x, f, new_xyz, rows=r, idx=idx)
        x = new_xyz
        data_dict[name + "_xyz"] = x
        data_dict[name + "_features"] = f
        if name in ("sa1", "sa2"):
            data_dict[name + "_inds"] = inds
    main.wait_event(nn_ev)
    if not capturing:
        for t in nn1 + nn2:
            t.record_stream(main)
    feats = fp_forward(bb.fp1, data_dict["sa3_xyz"], data_dict["sa4_xyz"], data_dict["sa3_features"],
                       data_dict["sa4_features"], nn=nn1)
    if feats is None:
        feats = bb.fp1(data_dict["sa3_xyz"], data_dict["sa4_xyz"], data_dict["sa3_features"], data_dict["sa4_features"])
    # "_fp2_features_out": an optional caller-provided (B, 256, num_seed) fp32 buffer for the final features
    # (GraphedBackbone(pack=...) points it into the buffer that leaves the device in one copy)
    up = fp_forward(bb.fp2, data_dict["sa2_xyz"], data_dict["sa3_xyz"], data_dict["sa2_features"], feats, nn=nn2,
                    out=data_dict.get("_fp2_features_out"))
    feats = up if up is not None else bb.fp2(data_dict["sa2_xyz"], data_dict["sa3_xyz"], data_dict["sa2_features"], feats)
    data_dict["fp2_features"] = feats
    data_dict["fp2_xyz"] = data_dict["sa2_xyz"]
    num_seed = data_dict["fp2_xyz"].shape[1]
    # same values as the reference's view sa1_inds[:, :num_seed] (backbone_module.py:134), materialised:
    # a strided device->host copy of this tensor costs the serving loop 0.1 ms per step, a contiguous
    # one nothing (tools/e2e_probe.py)
    data_dict["fp2_inds"] = data_dict["sa1_inds"][:, 0:num_seed].contiguous()
    return data_dict


class _FpPlan:
    def __init__(self, module):
        self.layers = []
        for blk in module.mlp.children():
            w, b = _fold_block(blk)
            self.layers.append((tcm.PackedWeight(w), b))
        self.c_in = self.layers[0][0].k
        self.key = _version_key(module)

    @staticmethod
    def supported(module):
        blocks = list(module.mlp.children())
        if not blocks or not _plain_relu_mlp(module.mlp, len(blocks)):
            return False
        outs = [b.conv.out_channels for b in blocks]
        if any(o % 16 or o > 256 for o in outs):
            return False
        return all(o % 64 == 0 for o in outs[:-1])  # a hidden width is the next layer's K


def fp_neighbours(unknown, known):
    """(idx (B,n,3) i32, weight (B,n,3) f32) of an FP layer: three_nn + normalised 1/(d+1e-8) weights."""
    b, n = unknown.shape[0], unknown.shape[1]
    idx = torch.empty(b, n, 3, device=unknown.device, dtype=torch.int32)
    w = torch.empty(b, n, 3, device=unknown.device, dtype=torch.float32)
    L.call("pn2_three_nn_weights", b, n, known.shape[1], L.ptr(unknown.contiguous()), L.ptr(known.contiguous()),
           L.ptr(idx), L.ptr(w), L.stream_of(unknown), device=unknown.device)
    return idx, w


def fp_forward(module, unknown, known, unknow_feats, known_feats, nn=None, out=None):
    """PointnetFPModule.forward (pointnet2_modules.py:371-416) in 1 + n_layers launches: the fused
    three_nn/interpolate/concat front end, then one tcgen05 linear kernel per MLP layer (the last
    one stores straight into the reference's (B, C, n) layout)."""
    if module._fused is False:
        return None
    if module._fused is None or module._fused.key != _version_key(module):
        if not _FpPlan.supported(module):
            module._fused = False
            return None
        module._fused = _FpPlan(module)
    plan = module._fused
    b, n = unknown.shape[0], unknown.shape[1]
    m = known.shape[1]
    c1 = unknow_feats.shape[1] if unknow_feats is not None else 0
    c2 = known_feats.shape[1]
    if c1 + c2 != plan.c_in or known_feats.shape[2] != m or unknown.dtype != torch.float32:
        return None
    dev = unknown.device
    kpad = plan.layers[0][0].kpad
    x = torch.empty(b * n, kpad, device=dev, dtype=torch.bfloat16)
    uf = unknow_feats.contiguous().float() if unknow_feats is not None else None
    if nn is not None:   # neighbours / weights computed ahead of time (backbone_forward: on the sampling stream)
        L.call("pn2_fp_prepare_indexed", b, n, m, c1, c2, kpad, L.ptr(nn[0]), L.ptr(nn[1]), L.ptr(uf),
               L.ptr(known_feats.contiguous().float()), L.ptr(x), L.stream_of(unknown), device=dev)
    else:
        L.call("pn2_fp_prepare", b, n, m, c1, c2, kpad, L.ptr(unknown.contiguous()), L.ptr(known.contiguous()),
               L.ptr(uf), L.ptr(known_feats.contiguous().float()), L.ptr(x), L.stream_of(unknown), device=dev)
    last = len(plan.layers) - 1
    for i, (w, bias) in enumerate(plan.layers):
        if i < last:
            x = tcm.linear(x, w, bias, relu=True, out="bf16")
        else:
            x = tcm.linear(x, w, bias, relu=True, out="cm", rows_per_batch=n, y=out)
    return x
