"""CPU pin of the oracle against the UNMODIFIED reference CUDA extension: tests/golden/reference_ext.npz
holds what /root/reference/lib/pointnet2/_ext_src (compiled for sm_100a, run on a B200 by
tests/golden/make_golden_ref_ext.py) returned on the seeded inputs of tests/golden/ref_ext_cases.py.
Indices, distances and copied / interpolated features must be bit-identical; the two scatter-add
gradients (fp32 atomics, order not defined) within 1e-5."""
import os

import numpy as np
import pytest

import ref_ext_cases as cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_ext.npz"))
    assert "reference _ext" in str(g["_producer"]) and "B200" in str(g["_producer"])
    return g


@pytest.mark.parametrize("case", sorted(cases.fps_cases()))
def test_fps_oracle_equals_reference_extension(oracle, gold, case):
    x, m = cases.fps_cases()[case]
    got = oracle.ext.furthest_point_sampling(x, m).numpy()
    want = gold["fps_" + case]
    assert got.dtype == want.dtype == np.int32 and got.shape == want.shape
    assert np.array_equal(got, want)


def test_fps_keyed_formulation_equals_reference_extension(oracle, gold):
    # the associative (temp, key) arg-max every product kernel is built on
    for case in ("dups", "lattice", "origin", "n300", "n37"):
        x, m = cases.fps_cases()[case]
        assert np.array_equal(oracle.furthest_point_sampling_keyed(x, m).numpy(), gold["fps_" + case]), case


def test_ball_query_oracle_equals_reference_extension(oracle, gold):
    x, c, _ = cases.search_inputs()
    for r, ns in cases.BALL_CFG:
        want = gold["ball_%g_%d" % (r, ns)]
        got = oracle.ext.ball_query(c, x, r, ns).numpy()
        assert got.shape == want.shape == (2, 256, ns)
        assert np.array_equal(got, want), (r, ns)


def test_three_nn_oracle_equals_reference_extension(oracle, gold):
    _, c, u = cases.search_inputs()
    d2, idx = oracle.ext.three_nn(u, c)
    assert np.array_equal(idx.numpy(), gold["three_nn_idx"])
    assert np.array_equal(d2.numpy().view(np.uint32), gold["three_nn_dist2"].view(np.uint32))   # bit for bit


def test_index_ops_oracle_equals_reference_extension(oracle, gold):
    pts, idx, i3, w3, go = cases.index_inputs()
    gi = idx[:, :, 0].contiguous()
    assert np.array_equal(oracle.ext.group_points(pts, idx).numpy(), gold["group_points"])
    assert np.array_equal(oracle.ext.gather_points(pts, gi).numpy(), gold["gather_points"])
    got = oracle.ext.three_interpolate(pts, i3, w3).numpy()
    assert np.array_equal(got.view(np.uint32), gold["three_interpolate"].view(np.uint32))        # pinned fma order
    np.testing.assert_allclose(oracle.ext.group_points_grad(go, idx, 700).numpy(), gold["group_points_grad"],
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle.ext.gather_points_grad(go[..., 0].contiguous(), gi, 700).numpy(),
                               gold["gather_points_grad"], rtol=1e-5, atol=1e-5)
