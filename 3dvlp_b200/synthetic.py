"""Synthetic ScanNet-sized scenes for tests and bench.py (no dataset ships with the reference).

Recipe of SURVEY.md section 8(d): an axis-aligned room (floor, four walls) furnished with boxes,
5 mm surface jitter, centred so that some points fall inside the FPS near-origin skip
(|p|^2 <= 1e-3, sampling_gpu.cu:105-106), then subsampled to exactly N points the way the
reference dataset does it -- ``np.random.choice(N_raw, N, replace=N_raw < N)``
(lib/joint/dataset.py:609-612) -- so exact duplicates occur when N_raw < N.  Feature channels follow
the dataset row layout (scripts/preprocess/point_cloud.py:31-47, dataset.py:603-607):
``[x, y, z, (nx, ny, nz, multiview x128,) height]``.
"""
import numpy as np
import torch


def _box_surface(rs, n, lo, hi):
    """n points uniformly on the 6 faces of the box [lo, hi] + their outward normals."""
    size = hi - lo
    areas = np.array([size[1] * size[2], size[1] * size[2], size[0] * size[2], size[0] * size[2],
                      size[0] * size[1], size[0] * size[1]])
    face = rs.choice(6, size=n, p=areas / areas.sum())
    p = lo + rs.uniform(size=(n, 3)) * size
    nrm = np.zeros((n, 3))
    ax = face // 2
    side = face % 2
    p[np.arange(n), ax] = np.where(side == 0, lo[ax], hi[ax])
    nrm[np.arange(n), ax] = np.where(side == 0, -1.0, 1.0)
    return p, nrm


def make_scene(seed, n=40000, c_in=1):
    """One scene as a (n, 3 + c_in) float32 array; c_in in {0, 1, 4, 132}."""
    rs = np.random.RandomState(seed)
    L, W, H = rs.uniform(4, 9), rs.uniform(4, 9), rs.uniform(2.4, 3.2)
    n_raw = int(rs.randint(25000, 150001))
    n_floor = int(0.35 * n_raw)
    n_wall = int(0.35 * n_raw)
    n_obj = n_raw - n_floor - n_wall
    pts, nrm = [], []
    f = np.stack([rs.uniform(0, L, n_floor), rs.uniform(0, W, n_floor), np.zeros(n_floor)], 1)
    pts.append(f); nrm.append(np.tile([0.0, 0.0, 1.0], (n_floor, 1)))
    wall = rs.randint(0, 4, n_wall)
    u = rs.uniform(size=n_wall); z = rs.uniform(0, H, n_wall)
    wx = np.where(wall == 0, 0.0, np.where(wall == 1, L, u * L))
    wy = np.where(wall == 2, 0.0, np.where(wall == 3, W, u * W))
    wn = np.zeros((n_wall, 3))
    wn[wall == 0, 0] = 1; wn[wall == 1, 0] = -1; wn[wall == 2, 1] = 1; wn[wall == 3, 1] = -1
    pts.append(np.stack([wx, wy, z], 1)); nrm.append(wn)
    n_boxes = int(rs.randint(12, 31))
    per = np.full(n_boxes, n_obj // n_boxes); per[: n_obj % n_boxes] += 1
    for k in range(n_boxes):
        size = rs.uniform(0.3, 2.0, 3)
        size[2] = min(size[2], H)
        lo = np.array([rs.uniform(0, max(L - size[0], 1e-3)), rs.uniform(0, max(W - size[1], 1e-3)), 0.0])
        p, q = _box_surface(rs, int(per[k]), lo, lo + size)
        pts.append(p); nrm.append(q)
    p = np.concatenate(pts, 0) + rs.normal(0, 0.005, size=(n_raw, 3))
    q = np.concatenate(nrm, 0)
    p[:, 0] -= L / 2; p[:, 1] -= W / 2
    choice = rs.choice(n_raw, n, replace=n_raw < n)
    p, q = p[choice], q[choice]
    cols = [p]
    if c_in >= 4:
        cols.append(q)
    if c_in >= 132:
        cols.append(rs.uniform(0, 1, size=(n, 128)))
    if c_in >= 1:
        floor_height = np.percentile(p[:, 2], 0.99)
        cols.append((p[:, 2] - floor_height)[:, None])
    row = np.concatenate(cols, 1).astype(np.float32)
    assert row.shape[1] == 3 + c_in, "c_in must be one of 0, 1, 4, 132"
    return row


def make_batch(first_scene, batch, n=40000, c_in=1, pin=False):
    """(batch, n, 3 + c_in) float32 torch tensor; scene i uses seed 1000 + first_scene + i."""
    arr = np.stack([make_scene(1000 + first_scene + i, n, c_in) for i in range(batch)], 0)
    t = torch.from_numpy(arr)
    return t.pin_memory() if pin else t


def randomize_bn_stats(module, seed=7):
    """Non-trivial eval-mode BatchNorm: running mean ~ N(0, 0.1), var ~ U(0.5, 1.5) (section 8d)."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    return module


MAX_NUM_OBJ = 128   # lib/joint/dataset.py: padded ground-truth slots per scene
NUM_CLASS = 18


def make_labels(pc, seed, lang_num_max=8, max_obj=MAX_NUM_OBJ):
    """Synthetic ground truth for a (B, N, 3 + C) batch in the layout the reference dataset yields
    (lib/joint/dataset.py:780-830) and its losses read (lib/loss_helper/loss_detection.py, loss_grounding.py):
    5-20 axis-aligned objects per scene centred on scene points, padded to `max_obj` slots; per-point votes
    to the centre of the box a point lies in (vote_label repeats the single vote three times, mask 0/1);
    `lang_num_max` sentence slots per scene, each referring to one object."""
    rs = np.random.RandomState(seed)
    xyz = pc[..., :3].numpy()
    B, N = xyz.shape[:2]
    out = {
        "center_label": np.zeros((B, max_obj, 3), np.float32),
        "heading_class_label": np.zeros((B, max_obj), np.int64),
        "heading_residual_label": np.zeros((B, max_obj), np.float32),
        "size_class_label": np.zeros((B, max_obj), np.int64),
        "size_residual_label": np.zeros((B, max_obj, 3), np.float32),
        "sem_cls_label": np.zeros((B, max_obj), np.int64),
        "box_label_mask": np.zeros((B, max_obj), np.float32),
        "vote_label": np.zeros((B, N, 9), np.float32),
        "vote_label_mask": np.zeros((B, N), np.int64),
        "ref_center_label_list": np.zeros((B, lang_num_max, 3), np.float32),
        "ref_heading_class_label_list": np.zeros((B, lang_num_max), np.int64),
        "ref_heading_residual_label_list": np.zeros((B, lang_num_max), np.float32),
        "ref_size_class_label_list": np.zeros((B, lang_num_max), np.int64),
        "ref_size_residual_label_list": np.zeros((B, lang_num_max, 3), np.float32),
        "lang_num": np.zeros((B,), np.int64),
    }
    mean_size = mean_size_arr()
    for b in range(B):
        n_obj = int(rs.randint(5, 21))
        centre = xyz[b, rs.choice(N, n_obj, replace=False)] + rs.normal(scale=0.05, size=(n_obj, 3)).astype(np.float32)
        cls = rs.randint(0, NUM_CLASS, n_obj)
        size = np.clip(mean_size[cls] + rs.normal(scale=0.05, size=(n_obj, 3)), 0.2, None).astype(np.float32)
        out["center_label"][b, :n_obj] = centre
        out["size_class_label"][b, :n_obj] = cls
        out["size_residual_label"][b, :n_obj] = size - mean_size[cls]
        out["sem_cls_label"][b, :n_obj] = cls
        out["box_label_mask"][b, :n_obj] = 1
        for i in range(n_obj):
            inside = np.all(np.abs(xyz[b] - centre[i]) <= size[i] / 2, axis=1)
            out["vote_label"][b, inside] = np.tile(centre[i] - xyz[b, inside], (1, 3))
            out["vote_label_mask"][b, inside] = 1
        n_lang = int(rs.randint(1, lang_num_max + 1))
        ref = rs.randint(0, n_obj, n_lang)
        out["lang_num"][b] = n_lang
        out["ref_center_label_list"][b, :n_lang] = centre[ref]
        out["ref_size_class_label_list"][b, :n_lang] = cls[ref]
        out["ref_size_residual_label_list"][b, :n_lang] = size[ref] - mean_size[cls[ref]]
    return {k: torch.from_numpy(v) for k, v in out.items()}


def mean_size_arr():
    """(18, 3) class mean box sizes (the reference loads them from a dataset file,
    data/scannet/model_util_scannet.py:92; any fixed table serves a synthetic run)."""
    rs = np.random.RandomState(7)
    return rs.uniform(0.4, 1.6, size=(NUM_CLASS, 3)).astype(np.float32)
