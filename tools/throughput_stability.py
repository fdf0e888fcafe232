"""Repeat the device-resident throughput loop of bench.py inside ONE process to see whether the step
time is stable across repetitions.  Usage: python tools/throughput_stability.py [depth] [reps] [steps] [stagger_us]"""
import importlib, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 600
dev = torch.device("cuda:0")
torch.manual_seed(42)
bb = pkg.Pointnet2Backbone(input_feature_dim=1)
syn.randomize_bn_stats(bb)
bb = bb.eval().to(dev)
pool = [syn.make_batch(8 * i, 8, 40000, 1).to(dev) for i in range(8)]
graphs = [pkg.GraphedBackbone(bb, pool[0], throughput=depth >= 6) for _ in range(depth)]
lanes = [torch.cuda.Stream(device=dev) for _ in range(depth)]
main = torch.cuda.current_stream(dev)


def run(count):
    for st in lanes:
        st.wait_stream(main)
    t0 = time.perf_counter()
    for s in range(count):
        with torch.cuda.stream(lanes[s % depth]):
            graphs[s % depth](pool[s % len(pool)])
    cpu = time.perf_counter() - t0
    for st in lanes:
        main.wait_stream(st)
    return cpu


run(24)
torch.cuda.synchronize()
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cpu = run(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("rep %d: %.4f ms/step (%.0f scenes/s), cpu issue %.4f ms/step" % (r, ms / steps, 8 * steps / ms * 1e3, cpu * 1e3 / steps), flush=True)
