"""Host wrappers of the tcgen05 entry points (pn2_pack_weight_bf16, pn2_pack_rows_bf16,
pn2_linear_bf16 -- include/pointnet2_b200.h)."""
import torch

from . import _lib as L


def kpad64(k):
    return (k + 63) // 64 * 64


class PackedWeight:
    """A (n, k) fp32 weight as bf16 SWIZZLE_128B panels, K zero-padded to a multiple of 64."""

    def __init__(self, w):
        w = w.detach().contiguous().float()
        L.require_cuda_path(w)
        self.n, self.k = w.shape
        self.kpad = kpad64(self.k)
        self.data = torch.empty(self.n * self.kpad * 2, device=w.device, dtype=torch.uint8)
        L.call("pn2_pack_weight_bf16", self.n, self.k, self.kpad, L.ptr(w), L.ptr(self.data), L.stream_of(w),
               device=w.device)


def pack_rows(features, kpad=None):
    """(B, C, N) fp32 channel-major -> (B*N, kpad) bf16 point-major rows, zero padded."""
    L.check_contiguous(features, "features"); L.check_float(features, "features"); L.require_cuda_path(features)
    b, c, n = features.shape
    kpad = kpad or kpad64(c)
    out = torch.empty(b * n, kpad, device=features.device, dtype=torch.bfloat16)
    L.call("pn2_pack_rows_bf16", b, c, n, kpad, L.ptr(features), L.ptr(out), L.stream_of(features),
           device=features.device)
    return out


def pack_point_rows(pc, col0, c, kpad=None):
    """(B, N, ld) fp32 point rows, features in columns [col0, col0+c) -> (B*N, kpad) bf16 rows."""
    import ctypes
    L.check_contiguous(pc, "pc"); L.check_float(pc, "pc"); L.require_cuda_path(pc)
    b, n, ld = pc.shape
    kpad = kpad or kpad64(c)
    out = torch.empty(b * n, kpad, device=pc.device, dtype=torch.bfloat16)
    L.call("pn2_pack_point_rows_bf16", ctypes.c_longlong(b * n), ld, col0, c, kpad, L.ptr(pc), L.ptr(out),
           L.stream_of(pc), device=pc.device)
    return out


def linear(x, w, bias=None, relu=False, out="f32", rows_per_batch=0, y=None, terms=1):
    """act(x @ W^T + bias) on tensor cores.  x (M, kpad) bf16; w PackedWeight; out in
    {"f32": (M,n) fp32, "bf16": (M,n) bf16, "cm": (M/rows_per_batch, n, rows_per_batch) fp32}.
    ``terms`` > 1 (training, mlp_train.py): x holds split-precision rows, each bf16 term stored once --
    (M, terms * kb) as [h | m] or [h | m | l] -- while w spells out the 3 or 6 K blocks of the leading products
    (w.kpad = 3 kb or 6 kb); the kernel maps a logical K block to the stored one."""
    L.require_cuda_path(x)
    nblk = {1: 1, 2: 3, 3: 6}[terms]
    kb = w.kpad // nblk
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or x.shape[1] != terms * kb or kb * nblk != w.kpad:
        raise RuntimeError("linear: x must be a contiguous bf16 (M, %d) tensor" % (terms * kb))
    m = x.shape[0]
    mode = {"f32": 0, "bf16": 1, "cm": 2}[out]
    if y is not None:  # caller-provided output (e.g. a view into a buffer that leaves the device in one copy)
        want = {0: ((m, w.n), torch.float32), 1: ((m, w.n), torch.bfloat16),
                2: ((m // max(rows_per_batch, 1), w.n, rows_per_batch), torch.float32)}[mode]
        if tuple(y.shape) != want[0] or y.dtype != want[1] or not y.is_contiguous() or y.device != x.device:
            raise RuntimeError("linear: y must be a contiguous %s tensor of shape %s" % (want[1], want[0]))
    elif mode == 0:
        y = torch.empty(m, w.n, device=x.device, dtype=torch.float32)
    elif mode == 1:
        y = torch.empty(m, w.n, device=x.device, dtype=torch.bfloat16)
    else:
        y = torch.empty(m // rows_per_batch, w.n, rows_per_batch, device=x.device, dtype=torch.float32)
    if terms == 1:
        L.call("pn2_linear_bf16", m, w.n, w.kpad, L.ptr(x), L.ptr(w.data), L.ptr(bias), int(bool(relu)), mode,
               int(rows_per_batch), L.ptr(y), L.stream_of(x), device=x.device)
    else:
        L.call("pn2_linear_split_bf16", m, w.n, kb, terms, L.ptr(x), L.ptr(w.data), L.ptr(bias), int(bool(relu)), mode,
               int(rows_per_batch), L.ptr(y), L.stream_of(x), device=x.device)
    return y


def poll_device_error(device):
    import ctypes
    flag = ctypes.c_int(0)
    L.call("pn2_poll_device_error", ctypes.byref(flag), L.stream_of(torch.empty(0, device=device)), device=device)
    return flag.value
