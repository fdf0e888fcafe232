"""Which part of the end-to-end step limits it?  Variants of bench.py's e2e loop:
 full = H2D + graph + D2H, noout = H2D + graph, noin = graph + D2H, none = graph only."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
depth, steps = 16, 600
dev = torch.device("cuda:0")
torch.manual_seed(42)
bb = pkg.Pointnet2Backbone(input_feature_dim=1)
syn.randomize_bn_stats(bb)
bb = bb.eval().to(dev)
host = [syn.make_batch(8 * i, 8, 40000, 1).pin_memory() for i in range(8)]
devp = [h.to(dev) for h in host]
graphs = [pkg.GraphedBackbone(bb, devp[0], throughput=True) for _ in range(depth)]
lanes = [torch.cuda.Stream(device=dev) for _ in range(depth)]
main = torch.cuda.current_stream(dev)
outs = ("fp2_features", "fp2_xyz", "fp2_inds")
d = graphs[0](devp[0])
out_host = [{k: torch.empty(d[k].shape, dtype=d[k].dtype).pin_memory() for k in outs} for _ in range(depth)]


def run(count, mode):
    for st in lanes:
        st.wait_stream(main)
    for s in range(count):
        k = s % depth
        with torch.cuda.stream(lanes[k]):
            r = graphs[k](host[s % 8] if mode in ("full", "noout") else devp[s % 8])
            if mode in ("full", "noin"):
                for key in outs:
                    out_host[k][key].copy_(r[key], non_blocking=True)
    for st in lanes:
        main.wait_stream(st)


for mode in ("none", "noout", "noin", "full", "none", "full"):
    run(200, mode); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(steps, mode); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("%-6s %.4f ms/step  %.0f scenes/s" % (mode, ms, 8e3 / ms), flush=True)
