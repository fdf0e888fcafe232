"""Run the product FPS a few times on synthetic scenes (target for ncu captures).
Usage: python tools/run_fps_once.py [n m b reps [both]] -- `both` runs the 16-CTA cluster kernel and then
the single-SM bucketed kernel (pn2_set_option("fps_bucket", 1)) in the same process."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
b = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
variants = (0, 1) if len(sys.argv) > 5 and sys.argv[5] == "both" else (pkg._lib.get_option("fps_bucket"),)
x = syn.make_batch(0, b, n, 0).cuda().contiguous()
for v in variants:
    pkg._lib.set_option("fps_bucket", v)
    for _ in range(reps):
        idx, nx = pkg._ext.furthest_point_sampling_gather(x, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); idx, nx = pkg._ext.furthest_point_sampling_gather(x, m); e1.record(); torch.cuda.synchronize()
    print("fps n=%d m=%d b=%d bucket=%d: %.3f ms" % (n, m, b, v, e0.elapsed_time(e1)), int(idx.sum()))
