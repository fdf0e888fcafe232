"""BASELINE.json config 5 (scaling sweep) and config 4 (132 input channels) on one GPU: CUDA-graph
replay of the backbone forward per (points, SA1 npoint, batch, C_in), one batch in flight (latency
mode) and four in flight.  Usage: python tools/sweep.py out.json"""
import importlib, json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
dev = torch.device("cuda:0")
rows = []


def run(n, npoint, b, c_in, depth):
    torch.manual_seed(42)
    bb = pkg.Pointnet2Backbone(input_feature_dim=c_in)
    syn.randomize_bn_stats(bb)
    bb = bb.eval().to(dev)
    bb.sa1.npoint = npoint
    # one generated scene per distinct seed, tiled to the batch (generation dominates otherwise)
    base = syn.make_batch(0, min(b, 4), n, c_in)
    pcs = [torch.cat([base] * (b // base.shape[0]), 0).roll(s, 0).contiguous().to(dev) for s in range(2)]
    throughput = depth >= 4 and 16384 <= n <= 131072
    graphs = [pkg.GraphedBackbone(bb, pcs[0], throughput=throughput) for _ in range(depth)]
    streams = [torch.cuda.Stream() for _ in range(depth)]
    steps = 24 if n <= 50000 else 12

    def go(k):
        for s in range(k):
            with torch.cuda.stream(streams[s % depth]):
                graphs[s % depth](pcs[s % 2])
    go(depth * 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams:
        st.wait_event(e0)
    go(steps)
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    row = {"points": n, "sa1_npoint": npoint, "batch": b, "c_in": c_in, "in_flight": depth, "ms_per_step": round(ms, 3),
           "scenes_per_s": round(b / ms * 1e3, 1), "graph": all(g.graph is not None for g in graphs)}
    rows.append(row)
    print(json.dumps(row), flush=True)
    del graphs


for n in (20000, 50000, 100000):
    for npoint in (2048, 4096):
        for b in (4, 8, 16, 32):
            for depth in (1, 4):
                if n * b * depth > 100000 * 32 * 2:
                    continue
                run(n, npoint, b, 1, depth)
run(40000, 2048, 8, 132, 1)
run(40000, 2048, 8, 132, 4)
run(40000, 2048, 8, 1, 1)
run(40000, 2048, 8, 1, 4)
if len(sys.argv) > 1:
    json.dump({"device": torch.cuda.get_device_name(0), "rows": rows}, open(sys.argv[1], "w"), indent=1)
