"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/pointnet2_b200.h declares, and the host layer mirrors the reference's interface."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names += re.findall(r"\b(pn2_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    lib.pn2_abi_version.restype = ctypes.c_int
    assert lib.pn2_abi_version() >= 1


def test_nine_ext_functions_present(pkg):
    """bindings.cpp:11-24 of the reference."""
    for name in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn",
                 "three_interpolate", "three_interpolate_grad", "ball_query", "group_points",
                 "group_points_grad"):
        assert callable(getattr(pkg._ext, name))


def test_cpu_tensors_are_rejected_like_the_reference(pkg):
    u = pkg.pointnet2_utils
    with pytest.raises(RuntimeError, match="CPU not supported"):
        u.furthest_point_sample(torch.rand(1, 8, 3), 4)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        u.ball_query(0.2, 4, torch.rand(1, 8, 3), torch.rand(1, 2, 3))
    with pytest.raises(RuntimeError, match="must be a float tensor"):
        pkg._ext.furthest_point_sampling(torch.rand(1, 8, 3).double(), 4)
    with pytest.raises(RuntimeError, match="must be a contiguous tensor"):
        pkg._ext.furthest_point_sampling(torch.rand(1, 3, 8).transpose(1, 2), 4)
    with pytest.raises(RuntimeError, match="must be an int tensor"):
        pkg._ext.gather_points(torch.rand(1, 3, 8), torch.zeros(1, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CPU not supported"):
        pkg.nn_distance(torch.rand(1, 4, 3), torch.rand(1, 5, 3))


def test_state_dict_layout_matches_reference(pkg):
    """96 entries, 641 920 parameters at C_in=1 / 650 304 at C_in=132."""
    bb = pkg.Pointnet2Backbone(input_feature_dim=1)
    sd = bb.state_dict()
    assert len(sd) == 96
    assert sum(p.numel() for p in bb.parameters()) == 641920
    assert sum(p.numel() for p in pkg.Pointnet2Backbone(132).parameters()) == 650304
    for k in ("sa1.mlp_module.layer0.conv.weight", "sa1.mlp_module.layer0.bn.bn.running_mean",
              "sa4.mlp_module.layer2.bn.bn.num_batches_tracked", "fp1.mlp.layer0.conv.weight",
              "fp2.mlp.layer1.bn.bn.weight"):
        assert k in sd, k
    assert sd["sa1.mlp_module.layer0.conv.weight"].shape == (64, 4, 1, 1)
    assert sd["sa3.mlp_module.layer0.conv.weight"].shape == (128, 259, 1, 1)


def test_sa_ctor_mutates_mlp_like_reference(pkg):
    spec = [5, 16, 32]
    pkg.PointnetSAModuleVotes(npoint=4, radius=0.3, nsample=4, mlp=spec)
    assert spec[0] == 8  # pointnet2_modules.py:204-206


def test_install_dropin_registers_reference_import_paths(pkg):
    pkg.install_dropin()
    import importlib
    m = importlib.import_module("lib.pointnet2.pointnet2_modules")
    assert m.PointnetSAModuleVotes is pkg.PointnetSAModuleVotes
    ext = importlib.import_module("pointnet2._ext")
    assert ext is pkg._ext
    nd = importlib.import_module("utils.nn_distance")
    assert nd.nn_distance is pkg.nn_distance
    from lib.pointnet2.pytorch_utils import BNMomentumScheduler  # lib/joint/solver_3dvlp.py:21
    sched = BNMomentumScheduler(pkg.Pointnet2Backbone(1), lambda e: 0.5 * 0.5 ** e)
    sched.step(2)
    assert abs(pkg.Pointnet2Backbone  # noqa: B018
               and sched.model.sa1.mlp_module.layer0.bn.bn.momentum - 0.125) < 1e-12
    for k in ("lib", "lib.pointnet2", "utils", "pointnet2", "pointnet2._ext", "pointnet2_utils",
              "pytorch_utils", "pointnet2_modules", "lib.pointnet2.pointnet2_utils",
              "lib.pointnet2.pytorch_utils", "lib.pointnet2.pointnet2_modules", "utils.nn_distance"):
        sys.modules.pop(k, None)


def test_huber_loss_matches_formula(pkg):
    e = torch.linspace(-3, 3, 41)
    got = pkg.huber_loss(e, delta=0.7)
    a = e.abs()
    want = torch.where(a <= 0.7, 0.5 * a ** 2, 0.5 * 0.7 ** 2 + 0.7 * (a - 0.7))
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-7)


def test_options_are_process_wide_with_a_per_thread_override(pkg):
    """pn2_set_option / pn2_set_option_thread (include/pointnet2_b200.h): a thread that captures a CUDA graph
    under other settings must not change what other host threads launch."""
    import threading
    assert pkg.get_option("fps_bucket") == 0 and pkg.get_option("no_such_option") == -1
    seen = {}

    def other_thread():
        seen["before"] = pkg.get_option("fps_bucket")
        pkg.set_option_thread("fps_bucket", 1)          # override of THIS thread only
        seen["own"] = pkg.get_option("fps_bucket")

    with pkg.thread_options(fps_bucket=1, sa_tiles_per_group=16):
        assert pkg.get_option("fps_bucket") == 1 and pkg.get_option("sa_tiles_per_group") == 16
        t = threading.Thread(target=lambda: seen.setdefault("concurrent", pkg.get_option("fps_bucket")))
        t.start(); t.join()
    assert seen["concurrent"] == 0                        # the other thread saw the process-wide value
    assert pkg.get_option("fps_bucket") == 0 and pkg.get_option("sa_tiles_per_group") == 0
    t = threading.Thread(target=other_thread)
    t.start(); t.join()
    assert seen == {"concurrent": 0, "before": 0, "own": 1} and pkg.get_option("fps_bucket") == 0
    pkg.set_option("bq_grid", 0)
    try:
        assert pkg.get_option("bq_grid") == 0
    finally:
        pkg.set_option("bq_grid", 1)
    assert pkg.get_option("wgrad_tma") == 1   # tiled-TMA weight gradient is the default
    with pytest.raises(pkg._lib.Pn2Error):
        pkg.set_option("no_such_option", 1)
    with pytest.raises(pkg._lib.Pn2Error):
        pkg.set_option_thread("no_such_option", 1)


def test_launch_statistics_are_exposed(pkg):
    for name in ("fps_cluster_launches", "fps_bucket_launches", "bq_grid_launches", "bq_small_grid_launches",
                 "bq_scan_launches", "nn_small_launches", "nn_tiled_launches", "wgrad_tma_launches",
                 "wgrad_cpasync_launches"):
        assert pkg.get_stat(name) >= 0
    with pytest.raises(pkg._lib.Pn2Error):
        pkg.get_stat("no_such_statistic")


def test_training_path_selection_is_host_logic(pkg):
    """mlp_train.supported: which SharedMLPs the tcgen05 training path takes (CUDA fp32 input, conv1x1 without
    bias + BatchNorm + ReLU blocks, widths multiples of 64 up to 256); CPU tensors and other shapes fall to
    torch's layers, and the backend / precision switches are context managers."""
    MT = pkg.mlp_train
    ok = pkg.pytorch_utils.SharedMLP([8, 64, 128], bn=True)
    assert not MT.supported(ok, torch.rand(2, 8, 5, 4))                  # CPU tensor
    assert MT._PRECISION[0] == "bf16x6" and MT._BACKEND[0] == "tcgen05"
    with MT.precision("bf16"), MT.mlp_backend("torch"):
        assert MT._PRECISION[0] == "bf16" and MT._BACKEND[0] == "torch"
    assert MT._PRECISION[0] == "bf16x6" and MT._BACKEND[0] == "tcgen05"
    w = MT._W_BLOCKS
    r = MT._ROW_BLOCKS
    # the K-block pairing evaluates exactly the leading term products
    assert sorted(zip(r[2], w[2])) == [(0, 0), (0, 1), (1, 0)]
    assert sorted(zip(r[3], w[3])) == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0)]
