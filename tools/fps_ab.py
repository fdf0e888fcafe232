"""Time the SA1 furthest-point sampling call alone (8 scenes x 40000 -> 2048), both variants, CUDA events."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
b = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda:0")
x = syn.make_batch(0, b, n, 1)[..., :3].contiguous().to(dev)
outs = {}
for bucket in (1, 0):
    pkg.set_option("fps_bucket", bucket)
    for _ in range(3):
        out = pkg._ext.furthest_point_sampling_gather(x, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = pkg._ext.furthest_point_sampling_gather(x, m)
    e1.record(); torch.cuda.synchronize()
    outs[bucket] = out
    print("n=%d m=%d b=%d fps_bucket=%d: %.3f ms" % (n, m, b, bucket, e0.elapsed_time(e1) / 10))
same = all(torch.equal(u, v) for u, v in zip(outs[0], outs[1])) if isinstance(outs[0], (tuple, list)) else torch.equal(outs[0], outs[1])
print("bucketed == cluster (indices and coordinates, bit for bit):", same, "PN2_FB_VARIANT=" + os.environ.get("PN2_FB_VARIANT", "0"))
