"""Developer tool: the bench's TRAINING step eager vs captured as one CUDA graph (graph.GraphedTrainStep):
ms per step of both and the loss sequences of two identically initialised runs.
Usage: python tools/graph_train_step.py [steps]"""
import importlib, os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
LD = importlib.import_module("3dvlp_b200.loss_detection")
LG = importlib.import_module("3dvlp_b200.loss_grounding")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
msa = syn.mean_size_arr()
cfg = types.SimpleNamespace(num_class=18, num_heading_bin=1, num_size_cluster=18, mean_size_arr=msa,
                            mean_size_arr_tensor=torch.from_numpy(msa).to(dev))
host = syn.make_batch(0, 8, 40000, 1)
pc = host.to(dev)
labels = {k: v.to(dev) for k, v in syn.make_labels(host, seed=100, lang_num_max=8).items()}
cluster_ref = torch.randn(64, 256, generator=torch.Generator().manual_seed(300)).to(dev)


def make():
    torch.manual_seed(45)
    det = pkg.VoteNetDetector(input_feature_dim=1, mean_size_arr=msa).to(dev).train()
    opt = torch.optim.Adam(det.parameters(), lr=1e-3, capturable=True)

    def loss_fn():
        d = det({"point_clouds": pc})
        d.update(labels)
        d.update({"epoch": 10, "istrain": 1, "random": 0.7, "cluster_ref": cluster_ref})
        d = LD.get_detection_loss(d, cfg)
        d, ref_loss, _, _ = LG.compute_diou_loss(d, cfg)
        return d["vote_loss"] + 0.1 * d["objectness_loss"] + d["box_loss"] + 0.3 * ref_loss + d["diou_loss"]
    return det, opt, loss_fn


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = [fn() for _ in range(n)]
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, [float(v) for v in out]


W = 3
det, opt, loss_fn = make()
def eager():
    opt.zero_grad(set_to_none=True)
    loss = loss_fn(); loss.backward(); opt.step()
    return loss.detach().clone()
pre = [float(eager()) for _ in range(W)]
ms_e, le = timed(eager, steps)
print("eager  : %.2f ms/step; losses %s" % (ms_e, ["%.4f" % v for v in pre + le]))

det, opt, loss_fn = make()
g = pkg.GraphedTrainStep(loss_fn, opt, warmup=W)
ms_g, lg = timed(lambda: g.step().clone(), steps)
print("graphed: %.2f ms/step; losses (after %d eager warm-up steps) %s" % (ms_g, W, ["%.4f" % v for v in lg]))
rel = max(abs(a - b) / max(abs(a), 1e-6) for a, b in zip(le, lg))
print("max relative loss difference eager vs graphed over %d steps: %.2e" % (steps, rel))
flag = pkg.tensorcore.poll_device_error(dev) if hasattr(pkg, "tensorcore") else 0
print("device error flag:", flag, " peak memory %.2f GB" % (torch.cuda.max_memory_allocated(dev) / 2 ** 30))
