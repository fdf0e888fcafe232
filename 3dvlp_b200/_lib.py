"""ctypes binding of lib3dvlp_b200.so (C ABI: include/pointnet2_b200.h).

There is deliberately NO fallback: if the CUDA library has not been built, or a tensor lives on the
CPU, the call fails loudly.  (The reference has no CPU path either:
lib/pointnet2/_ext_src/src/sampling.cpp:38-40 ``TORCH_CHECK(false, "CPU not supported")``.)
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PN2_LIB_PATH") or os.path.join(_HERE, "lib3dvlp_b200.so")  # override: A/B builds of the kernels

_lib = None


class Pn2Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "3dvlp_b200: CUDA library %s is missing -- build it with `python 3dvlp_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.pn2_last_error.restype = ctypes.c_char_p
        L.pn2_abi_version.restype = ctypes.c_int
        L.pn2_get_stat.restype = ctypes.c_longlong
        _lib = L
    return _lib


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_of(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


# Kernel launches issued per C-ABI call (for bench.py's `gpu_launches` claim and per-op timing).
KERNELS_PER_CALL = {"pn2_nn_distance": 2, "pn2_nn_distance_grad": 2, "pn2_sa_group_mlp": 4}  # (pn2_ball_query on a large cloud launches 4; counted as 1)
launch_count = 0
profile = None  # when a list: every call appends (name, tag, start_event, end_event)
profile_tag = ""


def call(name, *args, device=None):
    """Invoke one C-ABI entry point on `device` (made current for the call) and raise on failure."""
    global launch_count
    L = lib()
    fn = getattr(L, name)
    launch_count += KERNELS_PER_CALL.get(name, 1)
    if profile is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        try:
            _call(L, fn, name, args, device)
        finally:
            ev1.record()
            profile.append((name, profile_tag, ev0, ev1))
        return
    _call(L, fn, name, args, device)


def _call(L, fn, name, args, device):
    if device is not None and device.index is not None and device.index != torch.cuda.current_device():
        with torch.cuda.device(device):
            rc = fn(*args)
    else:
        rc = fn(*args)
    if rc != 0:
        raise Pn2Error("%s failed (status %d): %s" % (name, rc, L.pn2_last_error().decode()))


def set_option(name, value):
    """pn2_set_option: "fps_bucket" (0 = low-latency cluster FPS, 1 = single-SM bucketed FPS for
    throughput with several batches in flight), "bq_grid" (hashed-grid ball query on/off),
    "sa_tiles_per_group" (0 = one fused-SA CTA per SM slot; k > 0 = fewer CTAs, >= k tiles per group),
    "wgrad_tma" (1 = training weight gradient fed by tiled TMA, 0 = cp.async copies)."""
    L = lib()
    if L.pn2_set_option(name.encode(), int(value)) != 0:
        raise Pn2Error(L.pn2_last_error().decode())


def get_option(name):
    return int(lib().pn2_get_option(name.encode()))


def set_option_thread(name, value):
    """pn2_set_option_thread: override an option for calls made by THIS host thread only
    (value=None or < 0 removes the override)."""
    L = lib()
    if L.pn2_set_option_thread(name.encode(), -1 if value is None else int(value)) != 0:
        raise Pn2Error(L.pn2_last_error().decode())


class thread_options:
    """``with thread_options(fps_bucket=1, sa_tiles_per_group=16): ...`` -- per-thread option overrides
    for the duration of the block; other host threads keep launching what the process-wide options say."""

    def __init__(self, **opts):
        self.opts = opts

    def __enter__(self):
        for k, v in self.opts.items():
            set_option_thread(k, v)
        return self

    def __exit__(self, *exc):
        for k in self.opts:
            set_option_thread(k, None)
        return False


def get_stat(name):
    """pn2_get_stat: how often a kernel variant was launched by this process (e.g. "fps_bucket_launches")."""
    v = int(lib().pn2_get_stat(name.encode()))
    if v < 0:
        raise Pn2Error("unknown statistic %r" % name)
    return v


# ---- argument checks with the reference's messages (include/utils.h:10-30) -----------------------

def check_contiguous(t, name):
    if not t.is_contiguous():
        raise RuntimeError(name + " must be a contiguous tensor")


def check_float(t, name):
    if t.dtype != torch.float32:
        raise RuntimeError(name + " must be a float tensor")


def check_int(t, name):
    if t.dtype != torch.int32:
        raise RuntimeError(name + " must be an int tensor")


def check_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(name + " must be a CUDA tensor")


def require_cuda_path(t):
    """The reference raises "CPU not supported" for CPU inputs; so do we."""
    if not t.is_cuda:
        raise RuntimeError("CPU not supported")
