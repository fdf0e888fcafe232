"""Golden vectors for TRAINING mode (batch-statistics BatchNorm, gradients), produced by the REFERENCE's own
modules in ``.train()`` mode, imported unmodified from /root/reference (build container only):

    python tests/golden/make_golden_train.py

* lib/pointnet2/pointnet2_modules.py PointnetSAModuleVotes (mlp 5+3 -> 64 -> 64 -> 128): forward, running
  statistics after the step, gradients w.r.t. xyz, features, every conv / BatchNorm parameter
* PointnetFPModule (mlp 20 -> 64 -> 128): forward, running statistics, gradients w.r.t. the skip features and
  the parameters
* models/base_module/backbone_module.py Pointnet2Backbone on a 3000-point scene: forward, a sample of
  parameter gradients and running statistics after one step
As in make_golden.py the nine CUDA ops are the C oracle (the reference has no CPU ops); everything above
them -- QueryAndGroup, SharedMLP with torch's own Conv2d / BatchNorm2d / max_pool2d in fp32, autograd --
is the reference's code.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from common import cloud, det_tensor, fill_state_dict  # noqa: E402
from make_golden import load_reference  # noqa: E402


def sa_inputs():
    pc = cloud(21, 2, 300, c_extra=5, extent=1.0, dup_frac=0.1, origin_pts=3)
    return pc[..., :3].contiguous(), pc[..., 3:].transpose(1, 2).contiguous()


def fp_inputs():
    unknown = cloud(31, 2, 50, extent=1.0)
    known = cloud(32, 2, 20, extent=1.0)
    uf = cloud(33, 2, 50, c_extra=5)[..., :8].transpose(1, 2).contiguous()   # (2, 8, 50)
    kf = cloud(34, 2, 20, c_extra=9)[..., :12].transpose(1, 2).contiguous()  # (2, 12, 20)
    return unknown, known, uf, kf


def backbone_scene():
    return cloud(41, 2, 3000, c_extra=1, extent=1.6, dup_frac=0.05, origin_pts=4)


BB_PARAMS = ("sa1.mlp_module.layer0.conv.weight", "sa1.mlp_module.layer2.bn.bn.weight", "sa2.mlp_module.layer1.conv.weight",
             "sa3.mlp_module.layer0.conv.weight", "sa4.mlp_module.layer2.bn.bn.bias", "fp1.mlp.layer0.conv.weight",
             "fp2.mlp.layer1.conv.weight", "fp2.mlp.layer1.bn.bn.weight")
BB_STATS = ("sa1.mlp_module.layer0.bn.bn.running_mean", "sa1.mlp_module.layer2.bn.bn.running_var",
            "sa3.mlp_module.layer1.bn.bn.running_var", "fp2.mlp.layer1.bn.bn.running_mean")


def main():
    torch.set_num_threads(4)
    ref_utils, ref_modules, RefBackbone, _ = load_reference()
    out = {}

    xyz, feats = sa_inputs()
    m = ref_modules.PointnetSAModuleVotes(npoint=40, radius=0.45, nsample=12, mlp=[5, 64, 64, 128], use_xyz=True,
                                          normalize_xyz=True)
    fill_state_dict(m)
    m.train()
    x = xyz.clone().requires_grad_(True)
    f = feats.clone().requires_grad_(True)
    nx, nf, ii = m(x, f)
    out["sa_new_xyz"], out["sa_features"], out["sa_inds"] = nx.detach().numpy(), nf.detach().numpy(), ii.numpy()
    ((nf * det_tensor("w_sa_f", tuple(nf.shape), 1.0)).sum() + (nx * det_tensor("w_sa_x", tuple(nx.shape), 1.0)).sum()).backward()
    out["sa_grad_xyz"], out["sa_grad_features"] = x.grad.numpy(), f.grad.numpy()
    for n, p in m.named_parameters():
        out["sa_grad_" + n] = p.grad.numpy()
    for n, b in m.named_buffers():
        out["sa_buf_" + n] = b.numpy()

    unknown, known, uf, kf = fp_inputs()
    fp = ref_modules.PointnetFPModule(mlp=[20, 64, 128])
    fill_state_dict(fp)
    fp.train()
    u = uf.clone().requires_grad_(True)
    o = fp(unknown, known, u, kf)
    out["fp_out"] = o.detach().numpy()
    (o * det_tensor("w_fp", tuple(o.shape), 1.0)).sum().backward()
    out["fp_grad_unknow_feats"] = u.grad.numpy()
    for n, p in fp.named_parameters():
        out["fp_grad_" + n] = p.grad.numpy()
    for n, b in fp.named_buffers():
        out["fp_buf_" + n] = b.numpy()

    bb = RefBackbone(input_feature_dim=1)
    fill_state_dict(bb)
    bb.train()
    d = bb({"point_clouds": backbone_scene()})
    out["bb_fp2_features_sub"] = d["fp2_features"][:, ::8, ::16].detach().numpy()
    out["bb_sa2_features_sub"] = d["sa2_features"][:, ::8, ::16].detach().numpy()
    out["bb_fp2_inds"] = d["fp2_inds"].numpy()
    (d["fp2_features"] * det_tensor("w_bb", tuple(d["fp2_features"].shape), 1.0)).sum().backward()
    named = dict(bb.named_parameters())
    for n in BB_PARAMS:
        out["bb_grad_" + n] = named[n].grad.numpy()
    sd = bb.state_dict()
    for n in BB_STATS:
        out["bb_buf_" + n] = sd[n].numpy()

    path = os.path.join(HERE, "reference_train.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
