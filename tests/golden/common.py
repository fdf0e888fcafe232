"""Deterministic inputs / parameters shared by make_golden.py (run once, in the build container,
against the reference's own Python layers) and by the tests that replay the fixtures.

Nothing here depends on a torch RNG: parameters are closed-form functions of their shape and name
so the generator and the tests build bit-identical state_dicts on any machine.
"""
import zlib

import numpy as np
import torch


def det_tensor(name, shape, scale=1.0, offset=0.0):
    """fp32 tensor in roughly [-scale, scale] + offset, a pure function of (name, shape)."""
    n = int(np.prod(shape)) if len(shape) else 1
    seed = zlib.crc32(name.encode()) % 9973
    i = np.arange(n, dtype=np.float64)
    v = np.sin(i * 0.61803398875 + seed * 0.137) * np.cos(i * 0.0173 + seed)
    return torch.from_numpy((v * scale + offset).astype(np.float32).reshape(shape))


def fill_state_dict(module):
    """Overwrite every parameter / buffer of `module` deterministically (conv weights ~ kaiming
    scale, BN affine near 1 / 0, non-trivial running statistics)."""
    sd = module.state_dict()
    new = {}
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            new[k] = torch.tensor(3, dtype=v.dtype)
        elif k.endswith("conv.weight") or k.endswith("fc.weight"):
            fan_in = int(np.prod(v.shape[1:]))
            new[k] = det_tensor(k, tuple(v.shape), scale=(2.0 / fan_in) ** 0.5 * 1.5)
        elif k.endswith("running_var"):
            new[k] = det_tensor(k, tuple(v.shape), scale=0.4, offset=1.0)
        elif k.endswith("running_mean"):
            new[k] = det_tensor(k, tuple(v.shape), scale=0.1)
        elif k.endswith("bn.weight"):
            new[k] = det_tensor(k, tuple(v.shape), scale=0.2, offset=1.0)
        else:  # biases
            new[k] = det_tensor(k, tuple(v.shape), scale=0.1)
    module.load_state_dict(new)
    return new


def cloud(seed, b, n, c_extra=0, extent=2.0, dup_frac=0.0, origin_pts=0):
    """(b, n, 3 + c_extra) fp32 point cloud from numpy's MT19937 (stable across versions).
    `dup_frac` of the points are exact duplicates of earlier ones (dataset sampling with
    replacement, lib/joint/dataset.py:609-612); `origin_pts` points sit inside |p|^2 <= 1e-3
    (the FPS skip of sampling_gpu.cu:105-106)."""
    rs = np.random.RandomState(seed)
    pc = rs.uniform(-extent, extent, size=(b, n, 3 + c_extra)).astype(np.float32)
    if c_extra:
        pc[..., 3:] = rs.uniform(0, 1, size=(b, n, c_extra)).astype(np.float32)
    nd = int(n * dup_frac)
    for i in range(b):
        if nd:
            src = rs.randint(0, n, size=nd)
            dst = rs.randint(0, n, size=nd)
            pc[i, dst] = pc[i, src]
        if origin_pts:
            where = rs.randint(0, n, size=origin_pts)
            pc[i, where, :3] = rs.uniform(-0.015, 0.015, size=(origin_pts, 3)).astype(np.float32)
    return torch.from_numpy(pc)


def fill_state_dict_conv1d(module):
    """Deterministic parameters for modules built from plain Conv1d / BatchNorm1d (VotingModule)."""
    sd = module.state_dict()
    new = {}
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            new[k] = torch.tensor(3, dtype=v.dtype)
        elif k.endswith("weight") and v.dim() >= 2:
            fan_in = int(np.prod(v.shape[1:]))
            new[k] = det_tensor(k, tuple(v.shape), scale=(2.0 / fan_in) ** 0.5 * 1.5)
        elif k.endswith("running_var"):
            new[k] = det_tensor(k, tuple(v.shape), scale=0.4, offset=1.0)
        elif k.endswith("running_mean"):
            new[k] = det_tensor(k, tuple(v.shape), scale=0.1)
        elif k.endswith("weight"):
            new[k] = det_tensor(k, tuple(v.shape), scale=0.2, offset=1.0)
        else:
            new[k] = det_tensor(k, tuple(v.shape), scale=0.1)
    module.load_state_dict(new)
    return new
