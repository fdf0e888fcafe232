"""Time the SA1-sized ball query (8 scenes x 40000 points, 2048 centres, r=0.2, ns=64) standalone."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
x = syn.make_batch(0, 8, 40000, 0).cuda().contiguous()
inds, c = pkg._ext.furthest_point_sampling_gather(x, 2048)
for _ in range(5):
    idx = pkg._ext.ball_query(c, x, 0.2, 64)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    idx = pkg._ext.ball_query(c, x, 0.2, 64)
e1.record(); torch.cuda.synchronize()
print("ball_query SA1: %.3f ms per call" % (e0.elapsed_time(e1) / 20), int(idx.sum()))
