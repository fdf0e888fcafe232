"""CPU tests of the oracle on the edge cases the reference kernels define and on
size-independent properties (adjointness of the three scatter-add gradients, FPS prefix / order
invariants).  The oracle is the checker of the `-m gpu` parity tests, so its own corner behaviour is
pinned here, without a GPU."""
import numpy as np
import pytest
import torch


def _cloud(seed, b, n, extent=2.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(b, n, 3, generator=g) - 0.5) * extent + 1.5   # away from the origin skip zone


# ---- furthest point sampling (sampling_gpu.cu:75-178) ----------------------------------------------

def test_fps_zero_samples_returns_empty(oracle):
    # sampling_gpu.cu:78 `if (m <= 0) return;` -- the (B, 0) output is simply untouched
    out = oracle.ext.furthest_point_sampling(_cloud(0, 2, 64), 0)
    assert out.shape == (2, 0) and out.dtype == torch.int32


def test_fps_first_index_is_zero_even_inside_the_skip_zone(oracle):
    # idxs[0] = 0 unconditionally (sampling_gpu.cu:90-91); points with |p|^2 <= 1e-3 are never chosen later
    xyz = _cloud(1, 1, 128)
    xyz[0, 0] = torch.tensor([0.01, 0.0, 0.01])
    xyz[0, 7] = torch.tensor([0.0, 0.02, 0.0])
    out = oracle.ext.furthest_point_sampling(xyz.contiguous(), 100)[0].tolist()
    assert out[0] == 0
    assert 7 not in out and 0 not in out[1:]


def test_fps_all_points_skipped_repeats_index_zero(oracle):
    # nobody eligible: best = -1 in every thread, the tree keeps besti = 0 (sampling_gpu.cu:95-115)
    xyz = torch.full((1, 40, 3), 0.001)
    out = oracle.ext.furthest_point_sampling(xyz, 9)
    assert out[0].tolist() == [0] * 9


@pytest.mark.parametrize("n", [1, 3, 37, 300, 513])
def test_fps_small_and_ragged_sizes_are_permutation_prefixes(oracle, n):
    # block size = opt_n_threads(n) < 512 below 512 points (cuda_utils.h:20-24): distinct points are
    # each picked once while any remains, and a shorter request is a prefix of a longer one
    xyz = _cloud(10 + n, 2, n)
    full = oracle.ext.furthest_point_sampling(xyz, n)
    for b in range(2):
        assert sorted(full[b].tolist()) == list(range(n))
    if n > 2:
        half = oracle.ext.furthest_point_sampling(xyz, n // 2)
        assert torch.equal(half, full[:, : n // 2])


def test_fps_more_samples_than_points_keeps_running(oracle):
    # m > n is legal in the reference: once every distance is 0 the arg-max falls back on the tie-break
    xyz = _cloud(3, 1, 5)
    out = oracle.ext.furthest_point_sampling(xyz, 9)[0].tolist()
    assert sorted(out[:5]) == [0, 1, 2, 3, 4]
    assert all(0 <= i < 5 for i in out[5:])


def test_fps_greedy_invariant(oracle):
    # size-independent property: sample j maximises the distance to the set {0..j-1} (up to ties)
    xyz = _cloud(4, 1, 700)
    m = 64
    inds = oracle.ext.furthest_point_sampling(xyz, m)[0].long()
    p = xyz[0].double()
    dmin = torch.full((700,), float("inf"), dtype=torch.float64)
    for j in range(1, m):
        dmin = torch.minimum(dmin, ((p - p[inds[j - 1]]) ** 2).sum(-1))
        assert dmin[inds[j]] >= dmin.max() * (1 - 1e-6)


def test_fps_keyed_equals_thread_tree_with_duplicates(oracle):
    # exact ties (duplicated points, as dataset.py:609-612 produces with replace=True)
    xyz = _cloud(5, 1, 600)
    xyz[0, 300:] = xyz[0, :300]
    xyz = xyz.contiguous()
    a = oracle.ext.furthest_point_sampling(xyz, 400)
    b = oracle.furthest_point_sampling_keyed(xyz, 400)
    assert torch.equal(a, b)


# ---- ball query (ball_query_gpu.cu:14-49) ----------------------------------------------------------

def test_ball_query_rows_are_ascending_then_first_hit_padding(oracle):
    xyz = _cloud(6, 2, 500, extent=1.0)
    centres = xyz[:, ::10].contiguous()
    ns = 12
    idx = oracle.ext.ball_query(centres, xyz, 0.15, ns)
    r2 = np.float32(0.15) * np.float32(0.15)
    for b in range(2):
        d2 = ((centres[b][:, None] - xyz[b][None]) ** 2).sum(-1)
        for j in range(centres.shape[1]):
            hits = torch.nonzero(d2[j] < float(r2) * (1 - 1e-5))[:, 0].tolist()  # safely-inside hits
            row = idx[b, j].tolist()
            k = len(set(row))
            assert row[:k] == sorted(row[:k])                  # ascending scan order
            assert all(v == row[0] for v in row[k:])           # padded with the first hit
            last = row[k - 1] if k == ns else xyz.shape[1]     # the scan stops after ns hits
            assert {h for h in hits if h <= last} <= set(row)  # no in-radius point was passed over
            assert j * 10 in row or k == ns                    # the centre is a hit unless ns came first


def test_ball_query_empty_inputs(oracle):
    xyz = _cloud(7, 1, 20)
    none = oracle.ext.ball_query(torch.zeros(1, 0, 3), xyz, 0.5, 4)
    assert none.shape == (1, 0, 4)
    far = oracle.ext.ball_query(torch.full((1, 3, 3), 100.0), xyz, 0.5, 4)
    assert int(far.abs().sum()) == 0                            # zero-initialised rows stay zero


# ---- gather / group / interpolate and their scatter-add gradients ----------------------------------

def test_gather_and_group_are_index_selects(oracle):
    g = torch.Generator().manual_seed(8)
    feats = torch.randn(2, 5, 40, generator=g)
    idx1 = torch.randint(0, 40, (2, 9), generator=g, dtype=torch.int32)
    idx2 = torch.randint(0, 40, (2, 9, 4), generator=g, dtype=torch.int32)
    got1 = oracle.ext.gather_points(feats, idx1)
    got2 = oracle.ext.group_points(feats, idx2)
    for b in range(2):
        assert torch.equal(got1[b], feats[b][:, idx1[b].long()])
        assert torch.equal(got2[b], feats[b][:, idx2[b].long()])


@pytest.mark.parametrize("op", ["gather", "group", "interpolate"])
def test_grad_is_the_adjoint_of_the_forward(oracle, op):
    # <fwd(x), y> == <x, bwd(y)> for every x, y: the property the full-size GPU tests also use
    g = torch.Generator().manual_seed(9)
    b, c, n, m, s = 2, 6, 50, 17, 5
    x = torch.randn(b, c, n, generator=g)
    if op == "gather":
        idx = torch.randint(0, n, (b, m), generator=g, dtype=torch.int32)
        idx[:, :4] = 3                                           # collisions: several outputs hit one input
        y = torch.randn(b, c, m, generator=g)
        fwd, bwd = oracle.ext.gather_points(x, idx), oracle.ext.gather_points_grad(y, idx, n)
    elif op == "group":
        idx = torch.randint(0, n, (b, m, s), generator=g, dtype=torch.int32)
        idx[:, 0] = 0
        y = torch.randn(b, c, m, s, generator=g)
        fwd, bwd = oracle.ext.group_points(x, idx), oracle.ext.group_points_grad(y, idx, n)
    else:
        idx = torch.randint(0, n, (b, m, 3), generator=g, dtype=torch.int32)
        w = torch.rand(b, m, 3, generator=g)
        w = w / w.sum(-1, keepdim=True)
        y = torch.randn(b, c, m, generator=g)
        fwd, bwd = oracle.ext.three_interpolate(x, idx, w), oracle.ext.three_interpolate_grad(y, idx, w, n)
    lhs = (fwd.double() * y.double()).sum()
    rhs = (x.double() * bwd.double()).sum()
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))


def test_grads_of_untouched_inputs_are_zero(oracle):
    idx = torch.full((1, 4), 2, dtype=torch.int32)
    gout = torch.ones(1, 3, 4)
    gin = oracle.ext.gather_points_grad(gout, idx, 6)
    assert gin[0, :, 2].tolist() == [4.0, 4.0, 4.0]
    gin[0, :, 2] = 0
    assert float(gin.abs().sum()) == 0.0


# ---- three_nn (interpolate_gpu.cu:14-64) -----------------------------------------------------------

def test_three_nn_matches_a_sorted_search(oracle):
    unknown, known = _cloud(11, 2, 70), _cloud(12, 2, 33)
    d2, idx = oracle.ext.three_nn(unknown, known)
    for b in range(2):
        full = ((unknown[b][:, None].double() - known[b][None].double()) ** 2).sum(-1)
        want = torch.sort(full, dim=1, stable=True)
        assert torch.equal(idx[b].long(), want.indices[:, :3])
        assert torch.allclose(d2[b].double(), want.values[:, :3], rtol=1e-5, atol=1e-7)
        assert bool((d2[b][:, 0] <= d2[b][:, 1]).all() and (d2[b][:, 1] <= d2[b][:, 2]).all())


def test_three_nn_single_known_point(oracle):
    d2, idx = oracle.ext.three_nn(_cloud(13, 1, 4), _cloud(14, 1, 1))
    assert idx.abs().sum() == 0
    assert bool(torch.isfinite(d2[..., 0]).all()) and bool(torch.isinf(d2[..., 1:]).all())


# ---- nn_distance (utils/nn_distance.py:32-59) ------------------------------------------------------

def test_nn_distance_is_symmetric_under_swapping_the_clouds(oracle):
    g = torch.Generator().manual_seed(15)
    a, b = torch.rand(3, 20, 3, generator=g), torch.rand(3, 31, 3, generator=g)
    d1, i1, d2, i2 = oracle.nn_distance(a, b)
    e1, j1, e2, j2 = oracle.nn_distance(b, a)
    assert torch.equal(d1, e2) and torch.equal(i1, j2) and torch.equal(d2, e1) and torch.equal(i2, j1)
    assert i1.dtype == torch.int64


def test_nn_distance_of_a_cloud_with_itself_is_zero(oracle):
    a = _cloud(16, 2, 25)
    d1, i1, d2, i2 = oracle.nn_distance(a, a)
    assert float(d1.abs().max()) == 0.0 and float(d2.abs().max()) == 0.0
    assert torch.equal(i1, torch.arange(25).expand(2, 25)) and torch.equal(i2, i1)
