"""Fused eval-mode paths for the SA / FP modules (filled in as the tcgen05 kernels land).

Returning None tells the module to use its generic op-by-op path.
"""


def sa_forward(module, xyz, features, inds):
    return None


def fp_forward(module, unknown, known, unknow_feats, known_feats):
    return None
