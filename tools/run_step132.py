"""Backbone forward with 132 input channels (BASELINE config 4): ncu target for the gather-mode SA1 kernel."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
torch.manual_seed(42)
bb = pkg.Pointnet2Backbone(input_feature_dim=132)
syn.randomize_bn_stats(bb)
bb = bb.eval().cuda()
pc = syn.make_batch(0, 8, 40000, 132).cuda()
with torch.no_grad():
    for s in range(3):
        d = bb({"point_clouds": pc})
torch.cuda.synchronize()
print("ok", float(d["fp2_features"].sum()))
