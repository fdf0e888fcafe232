"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own Python layers.

Run once in the build container (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

What is the reference here?  The reference's nine CUDA ops have no CPU path
(src/sampling.cpp:38-40), so its Python layers -- pointnet2_utils.py (QueryAndGroup, the autograd
wrappers), pointnet2_modules.py (PointnetSAModuleVotes, PointnetFPModule), pytorch_utils.py
(SharedMLP) and models/base_module/backbone_module.py (Pointnet2Backbone) -- are imported
UNMODIFIED from /root/reference with ``pointnet2_utils._ext`` pointed at the C oracle
(oracle/oracle.py: ext).  utils/nn_distance.py is pure torch and is run as is.  The fixtures
therefore pin (a) everything the reference does above the nine kernels, bit for bit in structure,
and (b) nn_distance including its own known-answer demo (nn_distance.py:95-122).  The nine kernels
themselves are pinned on the GPU box against the compiled reference extension
(tests/test_ref_ext_gpu.py).
"""
import builtins
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from common import cloud, fill_state_dict, fill_state_dict_conv1d  # noqa: E402
from oracle import oracle as O  # noqa: E402


def load_reference():
    builtins.__POINTNET2_SETUP__ = True  # pointnet2_utils.py:27 escape hatch (no compiled _ext here)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "lib", "pointnet2"))
    import pointnet2_utils as ref_utils
    ref_utils._ext = O.ext
    import pointnet2_modules as ref_modules
    from models.base_module.backbone_module import Pointnet2Backbone as RefBackbone
    from utils import nn_distance as ref_nn
    return ref_utils, ref_modules, RefBackbone, ref_nn


def main():
    torch.set_num_threads(1)
    ref_utils, ref_modules, RefBackbone, ref_nn = load_reference()
    out = {}

    # ---- nn_distance: the reference's own demo (seed 0) + the three modes on a second case -------
    np.random.seed(0)
    pc1 = torch.from_numpy(np.random.random((1, 5, 3)).astype(np.float32))
    pc2 = torch.from_numpy(np.random.random((1, 6, 3)).astype(np.float32))
    for tag, kw in (("l2", {}), ("huber", {"l1smooth": True}), ("l1", {"l1": True})):
        d1, i1, d2, i2 = ref_nn.nn_distance(pc1, pc2, **kw)
        out[f"nnd_demo_{tag}_d1"], out[f"nnd_demo_{tag}_i1"] = d1.numpy(), i1.numpy()
        out[f"nnd_demo_{tag}_d2"], out[f"nnd_demo_{tag}_i2"] = d2.numpy(), i2.numpy()
    a = cloud(11, 3, 37, extent=1.0)
    b = cloud(12, 3, 23, extent=1.0)
    b[:, 5] = 0.0; b[:, 6] = 0.0  # padded all-zero rows -> index ties (appendix C-6)
    for tag, kw in (("l2", {}), ("huber", {"l1smooth": True, "delta": 0.15}), ("l1", {"l1": True})):
        d1, i1, d2, i2 = ref_nn.nn_distance(a, b, **kw)
        out[f"nnd_b_{tag}_d1"], out[f"nnd_b_{tag}_i1"] = d1.numpy(), i1.numpy()
        out[f"nnd_b_{tag}_d2"], out[f"nnd_b_{tag}_i2"] = d2.numpy(), i2.numpy()

    # ---- QueryAndGroup --------------------------------------------------------------------------
    pc = cloud(21, 2, 300, c_extra=5, extent=1.0, dup_frac=0.1, origin_pts=3)
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
    inds = O.ext.furthest_point_sampling(xyz, 40)
    new_xyz = O.ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    qg = ref_utils.QueryAndGroup(0.45, 12, use_xyz=True, ret_grouped_xyz=True, normalize_xyz=True)
    nf, gx = qg(xyz, new_xyz, feats)
    out["qg_new_features"], out["qg_grouped_xyz"] = nf.numpy(), gx.numpy()

    # ---- PointnetSAModuleVotes (eval + train, max / avg / rbf, features None, user inds) ---------
    for tag, kw, use_feat in (("max", {}, True), ("avg", {"pooling": "avg"}, True),
                              ("rbf", {"pooling": "rbf"}, True), ("nofeat", {}, False)):
        m = ref_modules.PointnetSAModuleVotes(npoint=40, radius=0.45, nsample=12,
                                              mlp=[5 if use_feat else 0, 16, 16, 32], use_xyz=True,
                                              normalize_xyz=True, **kw)
        fill_state_dict(m)
        m.eval()
        with torch.no_grad():
            nx, nfeat, ii = m(xyz, feats if use_feat else None)
        out[f"sa_{tag}_new_xyz"], out[f"sa_{tag}_features"], out[f"sa_{tag}_inds"] = nx.numpy(), nfeat.numpy(), ii.numpy()
    m = ref_modules.PointnetSAModuleVotes(npoint=40, radius=0.45, nsample=12, mlp=[5, 16, 16, 32],
                                          use_xyz=True, normalize_xyz=False)
    fill_state_dict(m)
    m.train()
    user_inds = torch.arange(40, dtype=torch.int32).flip(0).repeat(2, 1).contiguous() * 3
    nx, nfeat, ii = m(xyz, feats, user_inds)
    out["sa_train_new_xyz"], out["sa_train_features"] = nx.detach().numpy(), nfeat.detach().numpy()
    sd = m.state_dict()
    out["sa_train_running_mean0"] = sd["mlp_module.layer0.bn.bn.running_mean"].numpy()
    out["sa_train_running_var0"] = sd["mlp_module.layer0.bn.bn.running_var"].numpy()

    # ---- PointnetFPModule (known / known=None) ---------------------------------------------------
    unknown = cloud(31, 2, 50, extent=1.0)
    known = cloud(32, 2, 20, extent=1.0)
    known[:, 3] = unknown[:, 7]  # exact coincidence: d = 0 -> weight ~ 1
    uf = cloud(33, 2, 50, c_extra=5)[..., :8].transpose(1, 2).contiguous()   # (2, 8, 50)
    kf = cloud(34, 2, 20, c_extra=9)[..., :12].transpose(1, 2).contiguous()  # (2, 12, 20)
    fp = ref_modules.PointnetFPModule(mlp=[20, 24, 16])
    fill_state_dict(fp)
    fp.eval()
    with torch.no_grad():
        out["fp_out"] = fp(unknown, known, uf, kf).numpy()
    fp2 = ref_modules.PointnetFPModule(mlp=[20, 24, 16])
    fill_state_dict(fp2)
    fp2.eval()
    with torch.no_grad():
        out["fp_known_none_out"] = fp2(unknown, None, uf, kf[:, :, :1].contiguous()).numpy()

    # ---- the whole backbone on a 3000-point scene (subsampled outputs keep the fixture small) -----
    bb = RefBackbone(input_feature_dim=1)
    fill_state_dict(bb)
    bb.eval()
    scene = cloud(41, 1, 3000, c_extra=1, extent=1.6, dup_frac=0.05, origin_pts=4)
    with torch.no_grad():
        d = bb({"point_clouds": scene})
    out["bb_sa1_inds"] = d["sa1_inds"].numpy()
    out["bb_sa2_inds"] = d["sa2_inds"].numpy()
    out["bb_fp2_inds"] = d["fp2_inds"].numpy()
    out["bb_sa4_xyz"] = d["sa4_xyz"].numpy()
    out["bb_sa1_features_sub"] = d["sa1_features"][:, ::4, ::32].numpy()
    out["bb_sa4_features_sub"] = d["sa4_features"][:, ::8, ::8].numpy()
    out["bb_fp2_features_sub"] = d["fp2_features"][:, ::8, ::16].numpy()
    out["bb_fp2_features_sum"] = np.array([d["fp2_features"].double().sum().item(),
                                           d["fp2_features"].double().abs().sum().item()])

    # ---- VotingModule (models/base_module/voting_module.py, pure torch) + the L2 norm of jointnet.py ---
    from models.base_module.voting_module import VotingModule as RefVoting
    vm = RefVoting(1, 64)
    fill_state_dict_conv1d(vm)
    vm.eval()
    sxyz = cloud(51, 2, 96, extent=1.0)[..., :3].contiguous()
    sfeat = cloud(52, 2, 96, c_extra=61)[..., :64].transpose(1, 2).contiguous()
    with torch.no_grad():
        vx, vf = vm(sxyz, sfeat)
    out["vote_xyz"], out["vote_features"] = vx.numpy(), vf.numpy()

    path = os.path.join(HERE, "reference_layers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
