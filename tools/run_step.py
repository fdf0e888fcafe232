"""One small backbone step through every kernel family of the hot path -- the target command for
`compute-sanitizer --tool memcheck|racecheck|synccheck` (profiles/README.md) and for quick ncu captures.

    python tools/run_step.py [points=20000] [scenes=2] [mode=latency|throughput|both]

latency: eager forward (16-CTA cluster FPS with DSMEM st.async exchange, hashed-grid ball query, fused SA with
red.global.max epilogue, FP front end, tcgen05 linears); throughput: the same under fps_bucket=1 /
sa_tiles_per_group=16 (bucketed single-SM FPS).  Also one nn_distance forward/backward and one DIoU launch."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "both"
dev = torch.device("cuda:0")
torch.manual_seed(42)
bb = pkg.Pointnet2Backbone(input_feature_dim=1)
syn.randomize_bn_stats(bb)
bb = bb.eval().to(dev)
pc = syn.make_batch(0, b, n, 1).to(dev)
outs = []
with torch.no_grad():
    if mode in ("latency", "both"):
        outs.append(bb({"point_clouds": pc})["fp2_features"].clone())
    if mode in ("throughput", "both"):
        with pkg.thread_options(fps_bucket=1, sa_tiles_per_group=16):
            outs.append(bb({"point_clouds": pc})["fp2_features"].clone())
a = torch.rand(4, 256, 3, device=dev, requires_grad=True)
c = torch.rand(4, 300, 3, device=dev)
d1, _, d2, _ = pkg.nn_distance(a, c)
(d1.sum() + d2.sum()).backward()
BU = importlib.import_module("3dvlp_b200.box_util")
iou, diou = BU.box3d_diou_pairs(torch.rand(2, 64, 3, device=dev), torch.rand(2, 64, 3, device=dev) + 0.1,
                                torch.rand(2, 5, 3, device=dev), torch.rand(2, 5, 3, device=dev) + 0.1)
torch.cuda.synchronize()
if len(outs) == 2:
    assert torch.equal(outs[0], outs[1])
flag = importlib.import_module("3dvlp_b200.tensorcore").poll_device_error(dev)
assert flag == 0, flag
print("ok", n, b, mode, float(outs[0].abs().mean()))
