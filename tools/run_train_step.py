"""Developer tool: the bench's TRAINING step (bench.train_step_leg's configuration) under torch.profiler --
device time per kernel over a few steps, to see where the 33 ms go.  Usage: python tools/run_train_step.py [steps]"""
import importlib, os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
LD = importlib.import_module("3dvlp_b200.loss_detection")
LG = importlib.import_module("3dvlp_b200.loss_grounding")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
msa = syn.mean_size_arr()
torch.manual_seed(45)
det = pkg.VoteNetDetector(input_feature_dim=1, mean_size_arr=msa).to(dev).train()
cfg = types.SimpleNamespace(num_class=18, num_heading_bin=1, num_size_cluster=18, mean_size_arr=msa,
                            mean_size_arr_tensor=torch.from_numpy(msa).to(dev))
opt = torch.optim.Adam(det.parameters(), lr=1e-3)
host = syn.make_batch(0, 8, 40000, 1)
pc = host.to(dev)
labels = {k: v.to(dev) for k, v in syn.make_labels(host, seed=100, lang_num_max=8).items()}
cluster_ref = torch.randn(64, 256, generator=torch.Generator().manual_seed(300)).to(dev)


def step():
    opt.zero_grad(set_to_none=True)
    d = det({"point_clouds": pc})
    d.update(labels)
    d.update({"epoch": 10, "istrain": 1, "random": 0.7, "cluster_ref": cluster_ref})
    d = LD.get_detection_loss(d, cfg)
    d, ref_loss, _, _ = LG.compute_diou_loss(d, cfg)
    loss = d["vote_loss"] + 0.1 * d["objectness_loss"] + d["box_loss"] + 0.3 * ref_loss + d["diou_loss"]
    loss.backward()
    opt.step()
    return loss.detach()


for _ in range(2):
    step()
torch.cuda.synchronize()
# host-side phases, each closed by a device synchronisation (so the sum exceeds the pipelined step)
import time
def phase_times():
    t = {}
    def tick(name, t0):
        torch.cuda.synchronize(); t[name] = (time.perf_counter() - t0) * 1e3; return time.perf_counter()
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    d = det({"point_clouds": pc}); t0 = tick("forward (detector)", t0)
    d.update(labels); d.update({"epoch": 10, "istrain": 1, "random": 0.7, "cluster_ref": cluster_ref})
    d = LD.get_detection_loss(d, cfg); d, ref_loss, _, _ = LG.compute_diou_loss(d, cfg)
    loss = d["vote_loss"] + 0.1 * d["objectness_loss"] + d["box_loss"] + 0.3 * ref_loss + d["diou_loss"]; t0 = tick("losses", t0)
    loss.backward(); t0 = tick("backward", t0)
    opt.step(); t0 = tick("optimizer", t0)
    return t
pt = phase_times(); pt = phase_times()
print("phases (ms, each synchronised):", {k: round(v, 2) for k, v in pt.items()})
# the same phases WITHOUT synchronisation: host issue time only
def issue_times():
    t = {}
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    d = det({"point_clouds": pc}); t["forward (detector)"] = (time.perf_counter() - t0) * 1e3; t0 = time.perf_counter()
    d.update(labels); d.update({"epoch": 10, "istrain": 1, "random": 0.7, "cluster_ref": cluster_ref})
    d = LD.get_detection_loss(d, cfg); d, ref_loss, _, _ = LG.compute_diou_loss(d, cfg)
    loss = d["vote_loss"] + 0.1 * d["objectness_loss"] + d["box_loss"] + 0.3 * ref_loss + d["diou_loss"]
    t["losses"] = (time.perf_counter() - t0) * 1e3; t0 = time.perf_counter()
    loss.backward(); t["backward"] = (time.perf_counter() - t0) * 1e3; t0 = time.perf_counter()
    opt.step(); t["optimizer"] = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    return t
it = issue_times(); it = issue_times()
print("host issue time per phase (ms):", {k: round(v, 2) for k, v in it.items()})
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
print("ms per step (no profiler): %.2f" % (e0.elapsed_time(e1) / steps))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if getattr(e, "device_time_total", 0) > 0]
tot = sum(e.device_time_total for e in ev)
print("device time per step: %.2f ms over %d kernels/step" % (tot / steps / 1e3, sum(e.count for e in ev) // steps))
for e in sorted(ev, key=lambda e: -e.device_time_total)[:40]:
    print("%8.1f us/step %5.1f%% x%-4d %s" % (e.device_time_total / steps, 100 * e.device_time_total / tot, e.count // steps, e.key[:110]))
