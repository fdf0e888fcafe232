"""Developer tool: the end-to-end leg of bench.py alone -- bursts of K = 20 steps over 20 lanes with staggered lane
priorities, six repetitions -- and, with E2E_TIMELINE=1, one more burst with an event after every copy / replay:
when each lane's input arrived, its forward finished and its result reached the host.
Usage: [E2E_TIMELINE=1] python tools/e2e_burst.py"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
model = bench.make_model(torch, pkg, syn, dev)
host_pool, dev_pool = bench.build_pool(torch, syn, dev, 0)
depth, K, W = 20, 20, 5
least, greatest = torch.cuda.Stream.priority_range()
levels = least - greatest + 1
prio = lambda k: greatest + (k * levels) // depth
outs = ("fp2_features", "fp2_xyz", "fp2_inds")
main = torch.cuda.current_stream(dev)
for _ in range(1):
    graphs = [pkg.GraphedBackbone(model, dev_pool[0], throughput=True, priority=prio(k), pack=outs) for k in range(depth)]
    lanes = [torch.cuda.Stream(device=dev, priority=prio(k)) for k in range(depth)]
    out_host = [torch.empty(g.packed.numel(), dtype=torch.uint8).pin_memory() for g in graphs]

    def run(first, count):
        for st in lanes:
            st.wait_stream(main)
        for s in range(count):
            k = s % depth
            with torch.cuda.stream(lanes[k]):
                graphs[k](host_pool[(first + s) % bench.POOL])
                out_host[k].copy_(graphs[k].packed, non_blocking=True)
        for st in lanes:
            main.wait_stream(st)

    run(0, 3 * depth); torch.cuda.synchronize()
    res = []
    for rep in range(6):
        run(0, W); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(W, K); e1.record(); torch.cuda.synchronize()
        res.append(K * bench.SCENES_PER_GPU / (e0.elapsed_time(e1) * 1e-3))
    if os.environ.get("E2E_TIMELINE"):
        # one more burst with an event after every copy / replay: when each lane's input arrived, its forward
        # finished and its result reached the host, relative to the start of the burst
        run(0, W); torch.cuda.synchronize()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(depth)]
        t0 = torch.cuda.Event(enable_timing=True)
        for st in lanes:
            st.wait_stream(main)
        t0.record(main)
        for st in lanes:
            st.wait_stream(main)
        for k in range(depth):
            with torch.cuda.stream(lanes[k]):
                g = graphs[k]
                g.static_in.copy_(host_pool[k], non_blocking=True); ev[k][0].record()
                g.graph.replay(); ev[k][1].record()
                out_host[k].copy_(g.packed, non_blocking=True); ev[k][2].record()
        torch.cuda.synchronize()
        print("lane: input on device / forward done / result on host (ms after the burst started)")
        for k in range(depth):
            print("  %2d: %6.2f %6.2f %6.2f" % ((k,) + tuple(t0.elapsed_time(e) for e in ev[k])))
    graphs[0].check()
    print("e2e scenes/s at K=20: %s  median %.0f" % (" ".join("%.0f" % v for v in res), sorted(res)[len(res) // 2]))
    del graphs, lanes, out_host
    torch.cuda.synchronize()
