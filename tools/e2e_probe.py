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


copy_lanes = [torch.cuda.Stream(device=dev) for _ in range(depth)]
done = [torch.cuda.Event() for _ in range(depth)]
import time


def run(count, mode):
    for st in lanes:
        st.wait_stream(main)
    t0 = time.perf_counter()
    for s in range(count):
        k = s % depth
        with torch.cuda.stream(lanes[k]):
            if mode == "side" and s >= depth:
                lanes[k].wait_event(done[k])   # the previous outputs of this lane have left
            r = graphs[k](host[s % 8] if mode in ("full", "noout") else devp[s % 8])
            if mode in ("full", "noin"):
                for key in outs:
                    out_host[k][key].copy_(r[key], non_blocking=True)
            elif mode == "one":      # a single output copy (features only)
                out_host[k]["fp2_features"].copy_(r["fp2_features"], non_blocking=True)
            elif mode == "small":    # only the two small outputs
                out_host[k]["fp2_xyz"].copy_(r["fp2_xyz"], non_blocking=True)
                out_host[k]["fp2_inds"].copy_(r["fp2_inds"], non_blocking=True)
            elif mode == "xyz":
                out_host[k]["fp2_xyz"].copy_(r["fp2_xyz"], non_blocking=True)
            elif mode == "inds":
                out_host[k]["fp2_inds"].copy_(r["fp2_inds"], non_blocking=True)
            elif mode == "indsc":
                out_host[k]["fp2_inds"].copy_(r["sa2_inds"], non_blocking=True)   # contiguous tensor of the same size
            elif mode == "side":     # outputs leave on a per-lane copy stream
                ev = torch.cuda.Event(); ev.record()
                with torch.cuda.stream(copy_lanes[k]):
                    copy_lanes[k].wait_event(ev)
                    for key in outs:
                        out_host[k][key].copy_(r[key], non_blocking=True)
                    done[k].record()
    global cpu_ms
    cpu_ms = (time.perf_counter() - t0) * 1e3 / count
    for st in lanes + copy_lanes:
        main.wait_stream(st)


print("fp2_inds contiguous:", d["fp2_inds"].is_contiguous(), "fp2_xyz:", d["fp2_xyz"].is_contiguous())
for mode in ("none", "xyz", "inds", "indsc", "small", "full"):
    run(200, mode); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(steps, mode); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("%-6s %.4f ms/step  %.0f scenes/s   cpu issue %.4f ms/step" % (mode, ms, 8e3 / ms, cpu_ms), flush=True)
