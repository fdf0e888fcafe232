"""The kernels OFF the eval hot path, run once each at representative sizes -- the target command for their
`ncu --set full` captures (profiles/README.md): nn_distance at its live shapes (vote loss: B*1024 clouds of
1 x 3 points; objectness: 256 x 256) and at a large shape, three_nn standalone, the axis-aligned box IoU/DIoU
pairs kernel, and one training-mode SharedMLP + max-pool layer stack (tcgen05 forward / dgrad / wgrad,
BatchNorm statistics and backward kernels) at SA2's shape."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
MT = importlib.import_module("3dvlp_b200.mlp_train")
BU = importlib.import_module("3dvlp_b200.box_util")
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B = 8
# nn_distance: vote loss, objectness, large
votes = torch.rand(B * 1024, 1, 3, generator=g).to(dev).requires_grad_(True)
gt = torch.rand(B * 1024, 3, 3, generator=g).to(dev)
prop = torch.rand(B, 256, 3, generator=g).to(dev)
cent = torch.rand(B, 256, 3, generator=g).to(dev)
big1, big2 = torch.rand(B, 20000, 3, generator=g).to(dev), torch.rand(B, 20000, 3, generator=g).to(dev)
for _ in range(2):
    d1, _, d2, _ = pkg.nn_distance(votes, gt, l1=True)
    (d1.sum() + d2.sum()).backward()
    pkg.nn_distance(prop, cent)
    pkg.nn_distance(big1, big2)
# three_nn standalone
u = torch.rand(B, 20000, 3, generator=g).to(dev)
k = torch.rand(B, 2048, 3, generator=g).to(dev)
for _ in range(2):
    pkg._ext.three_nn(u, k)
# box IoU / DIoU: 256 proposals x 8 sentences per scene, forward + backward
pc = torch.rand(B, 256, 3, generator=g).to(dev).requires_grad_(True)
ps = (torch.rand(B, 256, 3, generator=g) + 0.2).to(dev).requires_grad_(True)
gc, gs = torch.rand(B, 8, 3, generator=g).to(dev), (torch.rand(B, 8, 3, generator=g) + 0.2).to(dev)
for _ in range(2):
    iou, diou = BU.box3d_diou_pairs(pc, ps, gc, gs)
    (iou.sum() + diou.sum()).backward()
# training-mode SA2-shaped layer stack: (B, 131, 1024, 32) -> 128 -> 128 -> 256, max over 32
mlp = pkg.pytorch_utils.SharedMLP([131, 128, 128, 256], bn=True).to(dev).train()
x = torch.randn(B, 131, 1024, 32, generator=g).to(dev).requires_grad_(True)
for _ in range(2):
    y = MT.shared_mlp_maxpool(mlp, x)
    y.sum().backward()
torch.cuda.synchronize()
print("ok")
