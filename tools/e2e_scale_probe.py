"""Why does the end-to-end figure scale worse than the device-resident one?  Run under torchrun with N ranks:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_scale_probe.py [bind]
Every rank reports (all ranks active at the same time, barriers in between):
  * where it runs: CPU affinity, NUMA node of its GPU, NUMA node(s) of its pinned buffers' CPUs
  * host<->device copy bandwidth of its pinned buffers, H2D, D2H and both at once
  * graph replays only, graph + H2D, graph + D2H, full end-to-end: ms/step on the device and host issue time
`bind`: before any pinned allocation, restrict the rank's CPU affinity to the cores of its GPU's NUMA node
(first-touch then places the pinned pages there)."""
import importlib, json, os, sys, time
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
bind = len(sys.argv) > 1 and sys.argv[1] == "bind"


def gpu_numa(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        p = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
        return int(open(p).read()), bus
    except Exception as ex:
        return None, repr(ex)[:80]


def node_cpus(node):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:
        return set()


info = {"rank": rank, "affinity_before": sorted(os.sched_getaffinity(0))}
node, bus = gpu_numa(local)
info["gpu_numa"], info["gpu_bus"] = node, bus
if bind and node is not None and node >= 0:
    want = node_cpus(node) & os.sched_getaffinity(0)
    if want:
        os.sched_setaffinity(0, want)
        info["bound_to"] = "%d cores of node %d" % (len(want), node)
    else:
        info["bound_to"] = "no allowed core on node %d (cpuset excludes it)" % node
info["affinity"] = "%d cores: %s.." % (len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4])
try:
    info["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
except Exception:
    pass

torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


# ---- copy bandwidth, all ranks at once ----
n = 64 << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1); h_out.fill_(1)
d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed_copy(which, reps=10):
    def fn():
        if which in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if which in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    fn(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        fn()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return (2 if which == "both" else 1) * n / ms / 1e6


for w in ("h2d", "d2h", "both"):
    info["gbs_" + w] = round(timed_copy(w), 1)
    barrier()

# ---- the bench loop in four variants ----
depth, steps = 20, 300
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
    t0 = time.perf_counter()
    for s in range(count):
        k = s % depth
        with torch.cuda.stream(lanes[k]):
            r = graphs[k](host[s % 8] if mode in ("full", "h2d") else devp[s % 8])
            if mode in ("full", "d2h"):
                for key in outs:
                    out_host[k][key].copy_(r[key], non_blocking=True)
    cpu = (time.perf_counter() - t0) * 1e3 / count
    for st in lanes:
        main.wait_stream(st)
    return cpu


for mode in ("none", "h2d", "d2h", "full"):
    run(100, mode); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cpu = run(steps, mode); e1.record(); torch.cuda.synchronize()
    info["ms_" + mode] = round(e0.elapsed_time(e1) / steps, 4)
    info["cpu_ms_" + mode] = round(cpu, 4)
    barrier()
info["load_avg"] = os.getloadavg()[0]
for r in range(world):
    if r == rank:
        print(json.dumps(info), flush=True)
    barrier()
if world > 1:
    dist.destroy_process_group()
