"""Run a few backbone forward steps on the bench workload (8 scenes x 40k points, C_in=1): the target
command for ncu launch lists and full captures (see profiles/README.md)."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(42)
bb = pkg.Pointnet2Backbone(input_feature_dim=1)
syn.randomize_bn_stats(bb)
bb = bb.eval().cuda()
pc = syn.make_batch(0, 8, 40000, 1).cuda()
with torch.no_grad():
    for s in range(steps):
        if s == steps - 1:
            torch.cuda.synchronize()
            torch.zeros(1, device="cuda").fill_(1234.0)  # marker launch: the last step follows
        d = bb({"point_clouds": pc})
torch.cuda.synchronize()
print("ok", float(d["fp2_features"].sum()))
