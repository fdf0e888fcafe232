"""SharedMLP (+ max-pool over nsample) with autograd on the sm_100a kernels: training-mode BatchNorm
(batch statistics, running-statistics update), backward, weight gradients.

Replaces, for ``.train()`` and for any call that needs gradients, what the reference runs through cuDNN:
``SharedMLP`` = per layer Conv2d 1x1 (no bias) -> BatchNorm2d -> ReLU (pytorch_utils.py:11-36,67-120) followed
by ``F.max_pool2d`` over the nsample axis (pointnet2_modules.py:251-257), on the materialised grouped tensor
(B, C, npoint, nsample).  Here every contraction (forward, dgrad, wgrad) is a tcgen05 kernel with bf16
operands and fp32 accumulation; BatchNorm statistics are accumulated in double; everything else is fused
elementwise / reduction kernels (csrc/mlp_train.cu, C ABI in include/pointnet2_b200.h).

Numerics contract (restated by oracle.shared_mlp_pool_mixed).  Every operand of a contraction (input,
weights, hidden activations, dY) is carried as up to three bf16 terms, v = h + m + l with h = bf16(v),
m = bf16(v - h), l = bf16(v - h - m), and a product a.b is evaluated on the tensor cores as the sum of its
leading term products, each exact in the fp32 accumulator:
  "bf16x6" (default)  3 terms, 6 products (hh, mh, hm, lh, hl, mm): ~2^-24 relative, i.e. fp32-level -- the
                      reference trains in fp32, and a stack of batch-normalised layers amplifies operand
                      rounding by |mean|/std per layer (measured on the backbone: 8e-5 at SA1 -> 0.2 at FP2
                      with 16-bit operands);
  "bf16x3"            2 terms, 3 products: ~2^-16 relative (60x tighter than the TF32 convolutions torch runs
                      the reference with by default), half the operand traffic of bf16x6;
  "bf16"              1 term: the eval path's contract.  Fast, but the gradient of a max-pool is discontinuous:
                      with 2^-9 operand rounding the arg-max of a few per cent of the (channel, centre) entries
                      differs from the fp32 reference's and moves their whole gradient to another neighbour --
                      15-35 % rms gradient error measured on random inputs although every kernel is exact
                      w.r.t. its contract.  Per layer: Y = X W^T kept in fp32; mean / biased variance of Y over all positions
(double accumulators); A = relu(Y * scale + shift) in fp32; output = max over nsample of the last A (fp32).
Backward: BatchNorm / ReLU / max-pool gradients in fp32, dW = dY^T X and dX = dY W as above.
"""
import ctypes

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib as L
from . import tensorcore as tcm

_BACKEND = ["tcgen05"]   # or "torch": the reference's cuDNN layers (fp32), kept for semantic cross-checks
_PRECISION = ["bf16x6"]  # or "bf16x3", "bf16"
_TERMS = {"bf16": 1, "bf16x3": 2, "bf16x6": 3}
# LOGICAL K-block order of the operand rows / of the weights per number of terms (0 = h, 1 = m, 2 = l): block i of
# the rows meets block i of the weight; (row term, weight term) pairs = the leading products.  The weights are stored
# with these blocks spelled out; the rows store each term ONCE ([h], [h | m], [h | m | l]) and the kernels map a
# logical block to the stored one (csrc/linear.cu kLinRowBlock, csrc/mlp_train.cu kWgProd*)
_ROW_BLOCKS = {1: (0,), 2: (0, 1, 0), 3: (0, 1, 0, 2, 0, 1)}
_W_BLOCKS = {1: (0,), 2: (0, 0, 1), 3: (0, 0, 1, 0, 2, 1)}


class precision:
    """``with precision("bf16"): ...`` -- operand precision of the training contractions (module docstring)."""

    def __init__(self, name):
        assert name in _TERMS
        self.name = name

    def __enter__(self):
        self.prev = _PRECISION[0]
        _PRECISION[0] = self.name
        return self

    def __exit__(self, *exc):
        _PRECISION[0] = self.prev
        return False


class mlp_backend:
    """``with mlp_backend("torch"): ...`` runs SharedMLP + pooling through torch's own layers (fp32, cuDNN)
    -- the reference's arithmetic, used by tests as an fp32 cross-check; the default "tcgen05" is this
    module's kernels."""

    def __init__(self, name):
        assert name in ("tcgen05", "torch")
        self.name = name

    def __enter__(self):
        self.prev = _BACKEND[0]
        _BACKEND[0] = self.name
        return self

    def __exit__(self, *exc):
        _BACKEND[0] = self.prev
        return False


def supported(mlp, x):
    """True if `mlp` is a SharedMLP of conv1x1(no bias) + BatchNorm + ReLU blocks this path can run."""
    if _BACKEND[0] != "tcgen05" or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4:
        return False
    blocks = list(mlp.children())
    if not blocks:
        return False
    for blk in blocks:
        names = [n for n, _ in blk.named_children()]
        if names != ["conv", "bn", "activation"] or not isinstance(blk.activation, nn.ReLU):
            return False
        conv = blk.conv
        if conv.bias is not None or tuple(conv.kernel_size) != (1, 1) or conv.out_channels % 64 or conv.out_channels > 256:
            return False
        if not isinstance(blk.bn.bn, nn.BatchNorm2d) or not blk.bn.bn.affine or not blk.bn.bn.track_running_stats:
            return False
    return x.shape[1] == blocks[0].conv.in_channels and x.shape[0] * x.shape[2] <= 2 ** 31 - 1


def _f(v):
    return ctypes.c_float(v)


def _pack_weight(w2d, kpad, terms):
    """(n, k) fp32 -> PackedWeight over the K blocks of `terms` (each of width kpad, zero padded)."""
    n, k = w2d.shape
    wp = w2d.new_zeros(n, kpad)
    wp[:, :k] = w2d
    if terms > 1:
        h = wp.to(torch.bfloat16).float()
        m = (wp - h).to(torch.bfloat16).float()
        parts = (h, m, wp - h - m)
        wp = torch.cat([parts[t] for t in _W_BLOCKS[terms]], dim=1)
    return tcm.PackedWeight(wp)


def _pack_input(x3, kpad, terms):
    """(B, C, n) fp32 channel-major -> operand rows (B*n, terms * kpad): each bf16 term stored once, [h | m | l]."""
    if terms == 1:
        return tcm.pack_rows(x3, kpad)
    b, c, n = x3.shape
    out = torch.empty(b * n, terms * kpad, device=x3.device, dtype=torch.bfloat16)
    L.call("pn2_pack_rows_split_bf16", b, c, n, kpad, terms, L.ptr(x3), L.ptr(out), L.stream_of(x3), device=x3.device)
    return out


def _ll(v):
    return ctypes.c_longlong(v)


class _SharedMLPMaxPool(Function):
    @staticmethod
    def forward(ctx, x, bns, training, *params):
        """x (B, C0, M, S) fp32; bns: the BatchNorm2d modules (running buffers, eps, momentum);
        params: (conv weight, bn weight, bn bias) per layer.  Returns (B, C_last, M) fp32."""
        dev = x.device
        B, C0, M, S = x.shape
        P = B * M * S
        n_layers = len(bns)
        st = L.stream_of(x)
        xs, ys, affine = [], [], []
        split = _TERMS[_PRECISION[0]]       # number of bf16 terms per operand
        kw = tcm.kpad64(C0)                                                           # width of one operand block
        X = _pack_input(x.reshape(B, C0, M * S).contiguous(), kw, split)              # (P, split * kw) bf16
        out = arg = None
        for i, bn in enumerate(bns):
            w, gamma, beta = params[3 * i], params[3 * i + 1], params[3 * i + 2]
            co = w.shape[0]
            wp = _pack_weight(w.detach().reshape(co, -1).float(), kw, split)
            Y = tcm.linear(X, wp, None, relu=False, out="f32", terms=split)           # (P, co) fp32
            aff = torch.empty(4, co, device=dev, dtype=torch.float32)                 # scale, shift, mean, invstd
            if training:
                stats = torch.zeros(2, co, device=dev, dtype=torch.float64)
                L.call("pn2_column_stats", _ll(P), co, co, L.ptr(Y), L.ptr(stats[0]), L.ptr(stats[1]), st, device=dev)
                bn.num_batches_tracked.add_(1)
                # torch.nn.BatchNorm: momentum None = cumulative moving average
                mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                L.call("pn2_bn_finalize", co, _ll(P), L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(gamma), L.ptr(beta), _f(bn.eps),
                       _f(mom), L.ptr(bn.running_mean), L.ptr(bn.running_var), L.ptr(aff[0]), L.ptr(aff[1]), L.ptr(aff[2]),
                       L.ptr(aff[3]), st, device=dev)
            else:
                L.call("pn2_bn_finalize", co, _ll(0), None, None, L.ptr(gamma), L.ptr(beta), _f(bn.eps), _f(0.0),
                       L.ptr(bn.running_mean), L.ptr(bn.running_var), L.ptr(aff[0]), L.ptr(aff[1]), L.ptr(aff[2]),
                       L.ptr(aff[3]), st, device=dev)
            xs.append(X); ys.append(Y); affine.append(aff)
            if i + 1 < n_layers:
                X = torch.empty(P, split * co, device=dev, dtype=torch.bfloat16)
                L.call("pn2_bn_relu_rows", _ll(P), co, co, L.ptr(Y), L.ptr(aff[0]), L.ptr(aff[1]), split, L.ptr(X), st,
                       device=dev)
                kw = co
            else:
                out = torch.empty(B, co, M, device=dev, dtype=torch.float32)
                arg = torch.empty(B * M, co, device=dev, dtype=torch.int32)   # point-major (group, channel)
                L.call("pn2_bn_relu_maxpool", B * M, S, co, co, M, L.ptr(Y), L.ptr(aff[0]), L.ptr(aff[1]), L.ptr(out),
                       L.ptr(arg), st, device=dev)
        ctx.save_for_backward(arg, *xs, *ys, *affine, *params)
        ctx.meta = (B, C0, M, S, n_layers, bool(training), split)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, C0, M, S, n, training, split = ctx.meta
        saved = ctx.saved_tensors
        arg = saved[0]
        xs, ys, affine, params = saved[1:1 + n], saved[1 + n:1 + 2 * n], saved[1 + 2 * n:1 + 3 * n], saved[1 + 3 * n:]
        dev = dout.device
        P = B * M * S
        st = L.stream_of(dout)
        dout = dout.float().transpose(1, 2).contiguous()   # (B, M, C): point-major like the arg-max
        grads = [None] * (3 * n)
        da, ld_da = None, 0
        grad_x = None
        for i in range(n - 1, -1, -1):
            w, gamma = params[3 * i], params[3 * i + 1]
            co = w.shape[0]
            ci = w.numel() // co
            Y, aff, X = ys[i], affine[i], xs[i]
            s = torch.zeros(2, co, device=dev, dtype=torch.float64)
            L.call("pn2_bn_bwd_stats", _ll(P), co, ld_da, co, L.ptr(da), L.ptr(dout if da is None else None),
                   L.ptr(arg if da is None else None), S, M, L.ptr(Y), L.ptr(aff[0]), L.ptr(aff[1]), L.ptr(aff[2]),
                   L.ptr(aff[3]), L.ptr(s[0]), L.ptr(s[1]), st, device=dev)
            grads[3 * i + 1] = s[1].float()   # d gamma = sum dYhat * xhat
            grads[3 * i + 2] = s[0].float()   # d beta  = sum dYhat
            blocks = split                                   # stored term blocks per row: [h | m | l]
            dY = torch.empty(P, blocks * co, device=dev, dtype=torch.bfloat16)
            L.call("pn2_bn_bwd_rows", _ll(P), co, ld_da, co, L.ptr(da), L.ptr(dout if da is None else None),
                   L.ptr(arg if da is None else None), S, M, L.ptr(Y), L.ptr(aff[0]), L.ptr(aff[1]), L.ptr(aff[2]),
                   L.ptr(aff[3]), L.ptr(gamma), L.ptr(s[0]), L.ptr(s[1]), _ll(P if training else 0), split, L.ptr(dY), st,
                   device=dev)
            dW = torch.zeros(co, ci, device=dev, dtype=torch.float32)
            kx = X.shape[1] // blocks                        # width of one operand block of X
            # dW = dY^T X: ONE launch for all leading (dY term, X term) products -- every term block of both
            # operands is read once and the products accumulate into one TMEM tile (csrc/mlp_train.cu)
            L.call("pn2_wgrad_split_bf16", _ll(P), co, ci, blocks * co, blocks * kx, L.ptr(dY), L.ptr(X), split,
                   L.ptr(dW), st, device=dev)
            grads[3 * i] = dW.view_as(w)
            if i == 0 and not ctx.needs_input_grad[0]:
                break
            # dX = dY W: the linear kernel with W^T (ci rows, co columns) as its weight; ci padded to 16 rows,
            # at most 256 output columns per launch
            wt = w.detach().reshape(co, ci).t().contiguous()                           # (ci, co)
            pieces = []
            for c0 in range(0, ci, 256):
                c1 = min(ci, c0 + 256)
                rows_pad = (c1 - c0 + 15) // 16 * 16
                blk = torch.zeros(rows_pad, co, device=dev, dtype=torch.float32)
                blk[:c1 - c0] = wt[c0:c1]
                pieces.append((c0, c1, rows_pad, tcm.linear(dY, _pack_weight(blk, co, split), None, relu=False, out="f32", terms=split)))
            if i > 0:
                assert len(pieces) == 1   # hidden widths are <= 256
                da, ld_da = pieces[0][3], pieces[0][2]
            else:
                grad_x = torch.empty(B, C0, M * S, device=dev, dtype=torch.float32)
                for c0, c1, rows_pad, d in pieces:
                    L.call("pn2_unpack_rows", B, c1 - c0, M * S, rows_pad, _ll(C0 * M * S), L.ptr(d),
                           ctypes.c_void_p(grad_x.data_ptr() + 4 * c0 * M * S), st, device=dev)
                grad_x = grad_x.view(B, C0, M, S)
        return (grad_x, None, None, *grads)


def shared_mlp_maxpool(mlp, x):
    """max over the last axis of ``mlp(x)``: x (B, C0, M, S) -> (B, C_last, M).  Uses the BatchNorm layers'
    batch statistics and updates their running buffers when ``mlp.training`` (torch semantics), the frozen
    running statistics otherwise."""
    blocks = list(mlp.children())
    bns = [blk.bn.bn for blk in blocks]
    params = []
    for blk in blocks:
        params += [blk.conv.weight, blk.bn.bn.weight, blk.bn.bn.bias]
    return _SharedMLPMaxPool.apply(x.contiguous(), bns, mlp.training, *params)
