"""Developer tool: does a fresh process start in the slow graph-launch mode (~1 ms per launch),
and for how long?  bench.py's device-resident loop (20 throughput-mode graphs on 20 streams), 30 chunks of 100 steps:
per chunk the wall time and the host's enqueue time.  NVML_MODE=clock|reasons|both additionally queries NVML from the
main thread every 0.5 ms while the GPU works through the enqueued chunk (does a query stall the GPU's work?).
Usage: [NVML_MODE=...] python tools/launch_stall_probe.py [chunks]"""
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
with torch.no_grad():
    model({"point_clouds": dev_pool[0]})
depth = 20
graphs = [pkg.GraphedBackbone(model, dev_pool[0], throughput=True) for k in range(depth)]
lanes = [torch.cuda.Stream(device=dev) for k in range(depth)]
main = torch.cuda.current_stream(dev)
mode = os.environ.get("NVML_MODE", "none")
nv = h = None
if mode != "none":
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
queries = 0
rows = []
for c in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    t0 = time.perf_counter()
    for st in lanes:
        st.wait_stream(main)
    for s in range(100):
        with torch.cuda.stream(lanes[s % depth]):
            graphs[s % depth](dev_pool[s % bench.POOL])
    for st in lanes:
        main.wait_stream(st)
    t1 = time.perf_counter()
    if nv is not None:
        end = torch.cuda.Event()
        end.record(main)
        while not end.query():
            if mode in ("clock", "both"):
                nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            if mode in ("reasons", "both"):
                nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            queries += 1
            time.sleep(0.0005)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    rows.append(((t2 - t0) * 1e3, (t1 - t0) * 1e3))
print("NVML_MODE=%s, %d query rounds" % (mode, queries))
print("chunk wall ms  :", " ".join("%.0f" % r[0] for r in rows))
print("host enqueue ms:", " ".join("%.0f" % r[1] for r in rows))
