#!/usr/bin/env python
"""bench.py -- scenes/s of Pointnet2Backbone forward on 40 000-point synthetic ScanNet-sized scenes.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     CPU arm: the reference path on host cores

A "step" is one eval-mode forward of the backbone over one batch of 8 scenes per GPU
(BASELINE.json configs[1]: 8 scenes x 40k points on one B200).  Rank 0 prints ONE JSON line.

  value        whole-job scenes/s with the inputs already resident in HBM (rotating pool of input
               batches larger than L2, so no step re-reads a cached input)
  e2e          same metric through the public module API with HOST buffers: every step copies its
               (8, 40000, 4) fp32 batch from pinned host memory and reads fp2_features / fp2_xyz /
               fp2_inds back to pinned host memory inside the timed region
  roofline     the dominant kernel (SA1 furthest-point sampling) + per-op table in roofline_ops
  cpu_baseline the CPU oracle (reference op semantics restated in C + the reference's torch layers)
               timed on this box's host cores on a bounded sample (rank 0, N = 1 only)
  clocks       SM clock / throttle reasons sampled through NVML by the main thread while the timed region is still
               running on the GPU (after its last enqueue, before its closing synchronisation)
  timed_region_attempts / host_issue_ms   W warm-up + exactly K timed steps per attempt; a region the host could not
               keep fed (enqueue > 35 % of its device time) is measured again, at most twice, all attempts listed
  grounding_step_forward / train_step (eager and as ONE CUDA graph) / votenet_* / ref_cuda_ext / numerics   extra legs
"""
import argparse
import gc
import importlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENES_PER_GPU = 8
N_POINTS = 40000
C_IN = 1
POOL = 32          # distinct input batches: 32 x 5.12 MB = 164 MB > 126 MB of L2
BASE_SCENES = 16   # distinct generated scenes; batches re-use them under seeded point permutations


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, in-process through NVML
    (nvidia_ml_py).  Spawning `nvidia-smi -lms` instead was measured to stall CUDA-graph launches for
    milliseconds at a time (its start-up and every poll take driver-wide locks): an intermittent 2x
    loss on a step that is one graph launch per 0.3 ms.  In-process NVML queries go through the same
    driver, so they are not issued from a background thread either: the main thread samples AFTER it has
    enqueued the region's K steps and BEFORE the closing synchronisation -- the GPU is still working
    through the region (checked with the closing event's query()), but no launch is left to be delayed.
    Falls back to one nvidia-smi query after the region if NVML is unavailable."""

    def __init__(self, index):
        self.index, self.rows, self.nv, self.h = index, [], None, None
        self.in_region = 0

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._sample()  # first use of the NVML entry points happens outside the timed region
            self.rows = []
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.rows.append((mhz, mask))

    def sample_until(self, end_event, period_s=0.002, max_samples=200):
        """Called between the last enqueue and the closing synchronisation: sample while `end_event` (recorded
        after the K steps) is still pending, i.e. while the GPU is inside the timed region."""
        if self.nv is None:
            return
        try:
            while len(self.rows) < max_samples:
                self._sample()
                if end_event.query():  # the region was over when (or just after) this sample was taken
                    break
                self.in_region += 1
                time.sleep(period_s)
        except Exception:
            pass

    def stop(self):
        if self.nv is None:
            return self._one_shot()
        rows = self.rows
        nv = self.nv
        names = (("hw_slowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                 ("hw_thermal_slowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                 ("sw_power_cap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
        reasons = sorted({n for _, mask in rows for n, bit in names if mask & bit})
        sm = [r[0] for r in rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "samples_before_region_end": self.in_region,
                "source": "nvml, main thread, after the last enqueue of the timed region and before its closing sync"}

    def _one_shot(self):
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            r = [c.strip() for c in out.split(",")]
            reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6])
                       if v.lower().startswith("active")]
            return {"sm_mhz": float(r[0]), "sm_max_mhz": float(r[1]), "reasons": reasons, "samples": 1,
                    "source": "nvidia-smi after the region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: before any pinned allocation, restrict this rank to the host cores of its GPU's NUMA node
    (where the cgroup's cpuset allows): its pinned buffers are then first-touched on that node and the
    host<->device copies of the end-to-end loop stay off the socket interconnect.  Returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        node = None
        for cand in (bus, bus[-12:]):
            p = "/sys/bus/pci/devices/%s/numa_node" % cand
            if os.path.exists(p):
                node = int(open(p).read())
                break
        if node is None or node < 0:
            return "gpu numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        if not want:
            return "gpu on node %d, none of the %d allowed cores is (cpuset): not bound" % (node, len(allowed))
        if want == allowed:
            return "gpu on node %d, all %d allowed cores are too" % (node, len(allowed))
        os.sched_setaffinity(0, want)
        return "bound to %d of %d allowed cores on node %d" % (len(want), len(allowed), node)
    except Exception as ex:
        return "not bound (%s)" % repr(ex)[:80]


def build_pool(torch, syn, dev, rank):
    """POOL batches of (8, 40000, 4): BASE_SCENES generated scenes re-used under seeded permutations of
    the point order (a different point order is a different FPS / ball-query problem)."""
    base = syn.make_batch(rank * BASE_SCENES, BASE_SCENES, N_POINTS, C_IN)  # (16, N, 4) host
    g = torch.Generator().manual_seed(1234 + rank)
    host = torch.empty(POOL, SCENES_PER_GPU, N_POINTS, 3 + C_IN, dtype=torch.float32).pin_memory()
    for i in range(POOL):
        for j in range(SCENES_PER_GPU):
            s = (i * SCENES_PER_GPU + j) % BASE_SCENES
            if i * SCENES_PER_GPU + j < BASE_SCENES:
                host[i, j] = base[s]
            else:
                host[i, j] = base[s][torch.randperm(N_POINTS, generator=g)]
    return host, host.to(dev)


def make_model(torch, pkg, syn, dev):
    torch.manual_seed(42)  # scripts/joint_scripts/train_3dvlp.py:782
    bb = pkg.Pointnet2Backbone(input_feature_dim=C_IN)
    syn.randomize_bn_stats(bb)
    return bb.eval().to(dev) if dev is not None else bb.eval()


def cpu_reference_arm(args, torch, pkg, syn):
    """--impl reference: the reference path on the host cores (no GPU work at all)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O.ext.threads = cores
    bb = make_model(torch, pkg, syn, None)
    sd = bb.state_dict()
    scenes = syn.make_batch(0, 1, N_POINTS, C_IN)
    with torch.no_grad():
        for _ in range(args.warmup):
            O.backbone(scenes, sd)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.backbone(scenes, sd)
        dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = "1 scene (of the batch of 8) per step, B=1, N=40000, C_in=1, fp32, all host threads"
    line = {"impl": "reference", "metric": "scenes/sec Pointnet2Backbone fwd (40k pts)", "value": v,
            "unit": "scenes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Pointnet2Backbone fwd, %d scenes/GPU x %d pts, C_in=%d, eval" % (
                           SCENES_PER_GPU, N_POINTS, C_IN), "scenes_per_gpu": SCENES_PER_GPU, "points": N_POINTS,
                       "c_in": C_IN, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "scenes/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _load_reference_stack():
    """The reference's OWN Python stack from baseline/_ref (laid out by oracle/build_ref.py, git-ignored):
    its pointnet2_utils / pointnet2_modules / pytorch_utils / backbone_module, bound to its OWN compiled
    extension (oracle/_ref/pointnet2_ref_ext.so registered under the name its Python layer imports,
    ``pointnet2._ext``, lib/pointnet2/pointnet2_utils.py:25-33).  Returns the stock Pointnet2Backbone
    class, or None when either half is missing."""
    import importlib.util
    import types
    so = os.path.join(ROOT, "oracle", "_ref", "pointnet2_ref_ext.so")
    base = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(so) or not os.path.exists(os.path.join(base, "models", "base_module", "backbone_module.py")):
        return None
    spec = importlib.util.spec_from_file_location("pointnet2_ref_ext", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    parent = types.ModuleType("pointnet2")
    parent.__path__ = []
    parent._ext = ref
    sys.modules["pointnet2"] = parent
    sys.modules["pointnet2._ext"] = ref
    saved_path = list(sys.path)
    shadowed = {k: sys.modules.pop(k) for k in list(sys.modules)
                if k in ("lib", "models", "utils", "pointnet2_utils", "pointnet2_modules", "pytorch_utils")
                or k.startswith(("lib.", "models.", "utils."))}
    try:
        sys.path.insert(0, base)
        sys.path.insert(0, os.path.join(base, "lib", "pointnet2"))
        mod = importlib.import_module("models.base_module.backbone_module")
        assert os.path.realpath(mod.__file__).startswith(os.path.realpath(base)), mod.__file__
        return mod.Pointnet2Backbone
    finally:
        sys.path[:] = saved_path
        for k, v in shadowed.items():
            sys.modules.setdefault(k, v)


def reference_cuda_ext_leg(torch, pkg, model, dev_pool, steps):
    """GPU baseline = the UNMODIFIED reference, stock code path: its own ``Pointnet2Backbone``
    (models/base_module/backbone_module.py:76-135) over its own pointnet2_modules / pointnet2_utils and
    its own compiled CUDA extension, with this run's weights loaded through ``load_state_dict`` (same 96
    keys) and the same input batches.  None of this repo's modules, kernels or host layer is on that path.
    This is the "reference CUDA-extension semantics" bar of the north star; a measured baseline only."""
    RefBackbone = _load_reference_stack()
    if RefBackbone is None:
        return None
    dev = dev_pool[0].device
    ref_model = RefBackbone(input_feature_dim=C_IN)
    ref_model.load_state_dict(model.state_dict(), strict=True)
    ref_model = ref_model.to(dev).eval()
    with torch.no_grad():
        for s in range(2):
            out = ref_model({"point_clouds": dev_pool[s % POOL]})
        torch.cuda.synchronize()
        # the stock stack and the product must agree on what they compute: indices exactly
        mine = model({"point_clouds": dev_pool[1 % POOL]})
        same_inds = bool(torch.equal(out["fp2_inds"].int(), mine["fp2_inds"].int()) and torch.equal(out["sa2_inds"], mine["sa2_inds"]))
        feat_err = float((out["fp2_features"] - mine["fp2_features"].float()).abs().max() / out["fp2_features"].abs().max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            ref_model({"point_clouds": dev_pool[(2 + s) % POOL]})
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": SCENES_PER_GPU / (ms * 1e-3), "unit": "scenes/s", "ms_per_step": ms, "steps": steps,
            "indices_equal_to_product": same_inds, "fp2_features_max_norm_diff": feat_err,
            "what": "UNMODIFIED reference: its Pointnet2Backbone + pointnet2_modules/utils (baseline/_ref) on its own "
                    "_ext (sm_100a build) + torch/cuDNN layers, same weights and inputs, 1 GPU"}


def grounding_leg(torch, dist, pkg, syn, model, dev, dev_pool, host_pool, world, rank, steps=20):
    """BASELINE configs[2]/[3] at N GPUs, forward: per rank 8 scenes -> detector (backbone + voting + proposal,
    one CUDA-graph replay) -> detection losses (vote / objectness / box / semantic class, on nn_distance) ->
    DIoU ("OID") reference loss -> OCC/OSC contrastive losses with the north star's ONE collective: the
    L2-normalised object (8 x 256, 128) and sentence (8 x 8, 128) embeddings of every rank all-gathered in a
    single NCCL call over NVLink.  The language branch is out of scope: sentence embeddings and the
    matching head's logits are fixed synthetic tensors.  Returns a dict for the bench line (rank 0 prints)."""
    import types
    import numpy as np
    LD = importlib.import_module("3dvlp_b200.loss_detection")
    LG = importlib.import_module("3dvlp_b200.loss_grounding")
    D = importlib.import_module("3dvlp_b200.dist")
    L_MAX, K, H = 8, 256, 128
    msa = syn.mean_size_arr()
    det = pkg.VoteNetDetector(input_feature_dim=C_IN, mean_size_arr=msa)
    det.backbone_net = model
    torch.manual_seed(43)
    syn.randomize_bn_stats(det.vgen); syn.randomize_bn_stats(det.proposal)
    det = det.eval().to(dev)
    cfg = types.SimpleNamespace(num_class=18, num_heading_bin=1, num_size_cluster=18, mean_size_arr=msa,
                                mean_size_arr_tensor=torch.from_numpy(msa).to(dev))
    torch.manual_seed(44)   # same projection weights on every rank (a data-parallel replica)
    con = pkg.ContrastModule(cfg).to(dev)
    labels = {k: v.to(dev) for k, v in syn.make_labels(host_pool[0], seed=100 + rank, lang_num_max=L_MAX).items()}
    g = torch.Generator().manual_seed(200 + rank)
    lang_emb = torch.randn(SCENES_PER_GPU * L_MAX, H, generator=g).to(dev)
    cluster_ref = torch.randn(SCENES_PER_GPU * L_MAX, K, generator=g).to(dev)
    graph = pkg.GraphedBackbone(det, dev_pool[0], throughput=False)

    def step():
        out = graph(dev_pool[0])
        d = dict(out)
        d.update(labels)
        d.update({"epoch": 60, "istrain": 0, "random": 0.7, "lang_emb": lang_emb, "cluster_ref": cluster_ref,
                  "bbox_feature": out["aggregated_vote_features"]})
        d = LD.get_detection_loss(d, cfg)
        d, ref_loss, _, _ = LG.compute_diou_loss(d, cfg)
        d = con(d, gather=world > 1)
        loss = d["vote_loss"] + 0.1 * d["objectness_loss"] + d["box_loss"] + 0.3 * ref_loss + d["diou_loss"] \
            + d["lang_con_loss"] + d["iou_con_loss"]
        return loss, d

    with torch.no_grad():
        for _ in range(3):
            loss, d = step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss, d = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        # the collective alone, and its result against a plain list all_gather
        obj = torch.nn.functional.normalize(con.pc_proj(d["bbox_feature"]), dim=-1).reshape(-1, H)
        sent = torch.nn.functional.normalize(con.text_proj(lang_emb), dim=-1)
        ag_us, equal = None, None
        if world > 1:
            for _ in range(5):
                D.gather_contrastive_embeddings(obj, sent, normalize=False)
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            for _ in range(50):
                all_obj, all_sent = D.gather_contrastive_embeddings(obj, sent, normalize=False)
            e1.record()
            torch.cuda.synchronize()
            ag_us = e0.elapsed_time(e1) / 50 * 1e3
            ref_o = [torch.empty_like(obj) for _ in range(world)]
            ref_s = [torch.empty_like(sent) for _ in range(world)]
            dist.all_gather(ref_o, obj.contiguous()); dist.all_gather(ref_s, sent.contiguous())
            equal = bool(torch.equal(all_obj, torch.cat(ref_o)) and torch.equal(all_sent, torch.cat(ref_s)))
        vals = torch.tensor([ms, ag_us or 0.0, float(loss), 1.0 if equal in (None, True) else 0.0], device=dev, dtype=torch.float64)
        if world > 1:
            mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            mn = vals.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            ms, ag_us, equal = float(mx[0]), float(mx[1]), bool(mn[3] > 0.5)
    per_rank_bytes = (obj.numel() + sent.numel()) * 4
    return {"value": SCENES_PER_GPU * world / (ms * 1e-3), "unit": "scenes/s", "ms_per_step": ms, "steps": steps,
            "loss_rank0": float(loss), "lang_con_loss_rank0": float(d["lang_con_loss"]),
            "all_gather": None if world == 1 else {
                "us": ag_us, "bytes_sent_per_rank": per_rank_bytes, "bytes_received_per_rank": per_rank_bytes * world,
                "tensors": "(%d,%d) object + (%d,%d) sentence embeddings, fp32, one all_gather_into_tensor" % (
                    obj.shape[0], H, sent.shape[0], H),
                "equals_reference_all_gather": equal, "nranks": world, "backend": "nccl"},
            "what": "forward of BASELINE configs[2]: detector graph replay + vote/objectness/box/sem-cls losses + DIoU "
                    "reference loss + OCC/OSC contrastive losses%s; eval-mode forward, one batch in flight, max over ranks" % (
                        " with the cross-GPU embedding all-gather" if world > 1 else " (single GPU: no collective)")}


def train_step_leg(torch, dist, pkg, syn, dev, host_pool, world, rank, steps=5):
    """BASELINE configs[2] as a TRAINING step, per rank 8 scenes x 40 000 points: VoteNetDetector in train() mode
    (batch-statistics BatchNorm) -> vote / objectness / box / semantic-class losses + DIoU reference loss ->
    backward -> Adam step.  Every shared MLP of the backbone and of the vote aggregation runs forward AND
    backward on this repo's kernels (mlp_train.py: tcgen05 GEMMs with split-bf16 operands, own BatchNorm /
    max-pool kernels); sampling, grouping, interpolation and nn_distance are the same kernels as in eval.  The
    Conv1d heads of VotingModule / ProposalModule (built for eval only) use torch layers.
    Replicas are independent (the reference trains with DataParallel; BatchNorm statistics stay per GPU): no
    gradient collective is timed."""
    import types
    LD = importlib.import_module("3dvlp_b200.loss_detection")
    LG = importlib.import_module("3dvlp_b200.loss_grounding")
    msa = syn.mean_size_arr()
    torch.manual_seed(45)
    det = pkg.VoteNetDetector(input_feature_dim=C_IN, mean_size_arr=msa).to(dev).train()
    cfg = types.SimpleNamespace(num_class=18, num_heading_bin=1, num_size_cluster=18, mean_size_arr=msa,
                                mean_size_arr_tensor=torch.from_numpy(msa).to(dev))
    opt = torch.optim.Adam(det.parameters(), lr=1e-3, capturable=True)
    pc = host_pool[0].to(dev)
    labels = {k: v.to(dev) for k, v in syn.make_labels(host_pool[0], seed=100 + rank, lang_num_max=8).items()}
    g = torch.Generator().manual_seed(300 + rank)
    cluster_ref = torch.randn(SCENES_PER_GPU * 8, 256, generator=g).to(dev)

    def loss_fn():
        d = det({"point_clouds": pc})
        d.update(labels)
        d.update({"epoch": 10, "istrain": 1, "random": 0.7, "cluster_ref": cluster_ref})
        d = LD.get_detection_loss(d, cfg)
        d, ref_loss, _, _ = LG.compute_diou_loss(d, cfg)
        return d["vote_loss"] + 0.1 * d["objectness_loss"] + d["box_loss"] + 0.3 * ref_loss + d["diou_loss"]

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn()
        loss.backward()
        opt.step()
        return loss.detach()

    losses = []
    for _ in range(2):
        losses.append(float(step()))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.reset_peak_memory_stats(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    losses.append(float(loss))
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    mlp = importlib.import_module("3dvlp_b200.mlp_train")
    # the same iteration captured ONCE as a CUDA graph (graph.GraphedTrainStep): ~3 900 launches become one, the
    # step is bound by its kernels instead of by the host's launch rate
    gs, why = None, ""
    try:
        gs = pkg.GraphedTrainStep(loss_fn, opt, warmup=1)
        for _ in range(2):
            gs.step()
    except Exception as ex:  # capture depends on the driver / torch build: the eager figure stands
        gs, why = None, repr(ex)[:200]
    torch.cuda.synchronize()
    ok = torch.tensor([1.0 if gs is not None else 0.0], device=dev, dtype=torch.float64)
    if world > 1:  # every rank replays or none does: the timed loop below contains collectives
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok[0]) > 0.5:
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(steps):
            gloss = gs.step()
        e1.record()
        torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([gms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            gms = float(t[0])
        graphed = {"value": SCENES_PER_GPU * world / (gms * 1e-3), "ms_per_step": gms, "loss_last_rank0": float(gloss),
                   "what": "the same iteration replayed as one CUDA graph (GraphedTrainStep)"}
    else:
        graphed = {"unavailable": why or "capture failed on another rank"}
    return {"value": SCENES_PER_GPU * world / (ms * 1e-3), "unit": "scenes/s", "ms_per_step": ms, "steps": steps,
            "loss_first_last_rank0": [losses[0], losses[-1]], "operand_precision": mlp._PRECISION[0],
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30, "graphed": graphed,
            "what": "TRAINING step of BASELINE configs[2]: detector forward in train() mode + detection / DIoU losses + "
                    "backward + Adam, eager, 8 scenes x 40000 pts per GPU; shared MLPs forward and backward on own "
                    "tcgen05 kernels"}


def algorithmic_work():
    """Algorithmic bytes (HBM-bound ops) or FLOPs (tensor ops) per C-ABI call of one backbone forward on
    8 scenes x 40 000 points, C_in = 1 -- SURVEY.md section 8(d) / BASELINE.md section 4, keyed by
    (entry point, n-th call).  FPS / ball query / 3-NN are latency- or ALU-bound by construction; their
    HBM fraction is reported for completeness (bound "hbm (latency-bound)")."""
    b = SCENES_PER_GPU
    sa = [(N_POINTS, 2048, 64, (3 + C_IN, 64, 64, 128)), (2048, 1024, 32, (131, 128, 128, 256)),
          (1024, 512, 16, (259, 128, 128, 256)), (512, 256, 16, (259, 128, 128, 256))]
    w = {}
    for i, (n, m, ns, ch) in enumerate(sa):
        w[("pn2_furthest_point_sampling_gather", i)] = ("hbm (latency-bound)", b * (12 * n + 16 * m))
        w[("pn2_ball_query", i)] = ("hbm (latency-bound)", b * (12 * n + 12 * m + 4 * m * ns))
        flops = 2 * b * m * ns * (ch[0] * ch[1] + ch[1] * ch[2] + ch[2] * ch[3])
        w[("pn2_sa_group_mlp", i)] = ("tensor", flops)
    # point-wise linears: SA2..SA4 per-point halves (K = 128 / 256 / 256), then FP1, FP2 (2 layers each)
    lin = [(b * 2048, 128, 128), (b * 1024, 128, 256), (b * 512, 128, 256),
           (b * 512, 256, 512), (b * 512, 256, 256), (b * 1024, 256, 512), (b * 1024, 256, 256)]
    for i, (m, n, k) in enumerate(lin):
        w[("pn2_linear_bf16", i)] = ("tensor", 2 * m * n * k)
    for i, (n, m) in enumerate([(512, 256), (1024, 512)]):
        w[("pn2_fp_prepare", i)] = ("hbm", b * (12 * (n + m) + 4 * 256 * m + 4 * 256 * n + 2 * 512 * n))
    return w


def settle(torch, run_steps_fn, chunk=100, max_chunks=12):
    """Untimed chunks of `chunk` steps until two consecutive chunks agree within 10 % (the first launches of freshly
    instantiated graphs are occasionally issued at ~1 ms each: driver-side first-use work)."""
    prev = None
    for i in range(max_chunks):
        t0 = time.perf_counter()
        run_steps_fn(chunk)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= 1 and prev is not None and abs(dt - prev) <= 0.1 * prev:
            break
        prev = dt


def per_op_profile(torch, L, model, batch, reps=5):
    """Per-C-ABI-call device time with CUDA events on the launching stream (one forward per rep)."""
    acc = {}
    for _ in range(reps):
        L.profile = []
        with torch.no_grad():
            model({"point_clouds": batch})
        torch.cuda.synchronize()
        seen = {}
        for name, tag, e0, e1 in L.profile:
            k = seen.get(name, 0)
            seen[name] = k + 1
            acc.setdefault((name, k), []).append(e0.elapsed_time(e1))
        L.profile = None
    return {k: statistics.median(v) for k, v in acc.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--depth", type=int, default=20, help="batches in flight (graph replays on separate streams)")
    ap.add_argument("--lane-priorities", type=int, default=-1,
                    help="end-to-end leg: 1 = the lanes (and the kernels of their graphs) get descending stream "
                         "priorities, so batches in flight complete one after the other and their device->host reads "
                         "overlap later compute; 0 = equal priorities (all batches of a wave complete together); "
                         "-1 (default) = 1 for a burst (steps < 4 x depth), 0 for a long run, where fair sharing "
                         "already overlaps the reads of earlier steps")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        import torch
        pkg = importlib.import_module("3dvlp_b200")
        syn = importlib.import_module("3dvlp_b200.synthetic")
        cpu_reference_arm(args, torch, pkg, syn)
        return

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("3dvlp_b200")
    syn = importlib.import_module("3dvlp_b200.synthetic")
    L = pkg._lib
    L.lib()  # fail loudly now if the CUDA library is missing

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: not bound"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)

    model = make_model(torch, pkg, syn, dev)
    host_pool, dev_pool = build_pool(torch, syn, dev, rank)
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fwd_eager(batch):
        with torch.no_grad():
            return model({"point_clouds": batch})

    # one eager forward: counts our kernel launches per step and fills every host-side cache
    L.launch_count = 0
    fwd_eager(dev_pool[0])
    torch.cuda.synchronize()
    launches_per_step = L.launch_count
    # the public fast path: the whole forward captured once as a CUDA graph, replayed per batch
    # `depth` batches in flight: FPS is a latency-bound serial chain that leaves most SMs idle, so two
    # graph replays on two streams overlap one batch's sampling chain with the other's feature path.
    depth = max(1, args.depth)
    throughput_mode = depth >= 6  # many batches in flight: 1-SM bucketed FPS; few: 16-SM cluster FPS
    try:
        least, greatest = torch.cuda.Stream.priority_range()
    except Exception:
        least, greatest = 0, -1
    stagger = args.lane_priorities == 1 or (args.lane_priorities < 0 and args.steps < 4 * depth)
    levels = least - greatest + 1 if stagger else 1

    def lane_priority(k):
        return greatest + (k * levels) // depth if levels > 1 else 0

    # device-resident leg: equal priorities (highest throughput: every sampling chain starts at once)
    graphs = [pkg.GraphedBackbone(model, dev_pool[0], throughput=throughput_mode) for k in range(depth)]
    lanes = [torch.cuda.Stream(device=dev) for k in range(depth)]
    graphed = graphs[0]
    execution = "cuda_graph" if graphed.graph is not None else "eager (graph capture failed: %s)" % graphed.capture_error
    main_stream = torch.cuda.current_stream(dev)

    def run_steps(first, count, step_fn, graphs=graphs, lanes=lanes):
        """Issue `count` steps round-robin over the `depth` lanes; returns after joining them."""
        for st in lanes:
            st.wait_stream(main_stream)
        for s in range(count):
            k = s % depth
            with torch.cuda.stream(lanes[k]):
                step_fn(graphs[k], first + s)
        for st in lanes:
            main_stream.wait_stream(st)

    # ---------------- device-resident throughput --------------------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    # Untimed settling before the W warm-up steps: the first few hundred graph launches of a process are
    # occasionally issued at ~1 ms each (driver-side first-use work; measured with
    # tools/throughput_stability.py), which W = 3..24 steps do not cover.  Chunks of 100 steps until two
    # consecutive chunks agree within 10 %, at most 20 chunks.
    prev = None
    for chunk in range(20):
        t0 = time.perf_counter()
        run_steps(0, 100, lambda g, s: g(dev_pool[s % POOL]))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if chunk >= 2 and prev is not None and abs(dt - prev) <= 0.1 * prev:
            break
        prev = dt
    # a cyclic garbage collection inside a 6 ms timed region (K = 20) would be a visible share of it: the heap
    # is collected before and the collector paused until both timed regions are over
    gc.disable()

    def timed_region(step_fn, graphs_, lanes_, sample_clocks):
        """W untimed warm-up steps, then exactly K steps between two events, barrier + synchronize on both sides.
        The region measures the GPU only if the host kept it fed: enqueueing K = 20 steps takes ~0.8 ms of the
        region's ~6 ms, but a host that stalls for a few milliseconds in that window (observed in ~1 of 10 fresh
        processes: a slow first-touch / a descheduled thread; the eager legs of the same process then run 3x
        slower too) starves the device.  A region whose enqueue took more than 35 % of its device time (on any
        rank) is measured again, at most twice; every attempt is reported in `timed_region_attempts`."""
        attempts = []
        for _ in range(3):
            run_steps(0, args.warmup, step_fn, graphs_, lanes_)
            barrier()
            gc.collect()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            t_issue = time.perf_counter()
            run_steps(args.warmup, args.steps, step_fn, graphs_, lanes_)
            a1.record()
            issue_ms = (time.perf_counter() - t_issue) * 1e3
            if sample_clocks:
                clocks.sample_until(a1)
            barrier()
            region_ms = a0.elapsed_time(a1)
            attempts.append({"ms": round(region_ms, 3), "host_issue_ms": round(issue_ms, 3)})
            starved = torch.tensor([1.0 if issue_ms > 0.35 * region_ms else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(starved, op=dist.ReduceOp.MAX)
            if float(starved[0]) == 0.0:
                break
        return region_ms, issue_ms, attempts

    ms, host_issue_ms, attempts_dev = timed_region(lambda g, s: g(dev_pool[s % POOL]), graphs, lanes, True)
    launches = launches_per_step * args.steps
    clk = clocks.stop()

    # ---------------- end to end through the public API with host buffers -------------------------
    # A burst of K ~ depth steps under equal priorities completes all its batches together, so all the
    # device->host reads queue up behind the last kernel; with descending lane priorities (kernels keep the
    # priority of the stream their graph was captured on) batches complete one after the other and the reads
    # of the early ones travel under the compute of the late ones (measured at K = 20: 16.1k -> 19.0k scenes/s)
    # The step's result (fp2_features / fp2_xyz / fp2_inds) is packed into one contiguous buffer inside the graph:
    # the host issues ONE device->host copy per step (three calls per step: input copy, replay, output copy).
    outs = ("fp2_features", "fp2_xyz", "fp2_inds")
    e_graphs = [pkg.GraphedBackbone(model, dev_pool[0], throughput=throughput_mode, priority=lane_priority(k), pack=outs)
                for k in range(depth)]
    e_lanes = [torch.cuda.Stream(device=dev, priority=lane_priority(k)) for k in range(depth)] if stagger else lanes
    out_host = [torch.empty(g.packed.numel(), dtype=torch.uint8).pin_memory() for g in e_graphs]
    lane_of = {id(g): i for i, g in enumerate(e_graphs)}

    def e2e_step(g, s):
        g(host_pool[s % POOL])  # pinned host -> static device input, then the graph
        out_host[lane_of[id(g)]].copy_(g.packed, non_blocking=True)

    run_steps(0, 3 * depth, e2e_step, e_graphs, e_lanes)   # untimed settling of the copy path (first use of the pinned buffers)
    torch.cuda.synchronize()
    ms_e2e, host_issue_ms_e2e, attempts_e2e = timed_region(e2e_step, e_graphs, e_lanes, False)
    gc.enable()
    graphs[0].check()  # raises if any kernel bailed out of a bounded wait during the timed regions
    h2d = dev_pool[0].numel() * 4
    d2h = sum(nbytes for _, _, nbytes, _, _ in e_graphs[0].pack_layout)
    # the host copy of the last step of lane 0 holds what the device computed (views into the packed buffer)
    last = e_graphs[0].packed_views(out_host[0])
    assert torch.equal(last["fp2_inds"], e_graphs[0].static_out["fp2_inds"].cpu())

    # max over ranks, on-device timings
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    try:  # every rank takes part (the leg contains the collective); rank 0 reports it
        grounding = grounding_leg(torch, dist, pkg, syn, model, dev, dev_pool, host_pool, world, rank)
    except Exception as ex:
        if world > 1:
            raise
        grounding = {"error": repr(ex)[:300]}
    try:
        train_step = train_step_leg(torch, dist, pkg, syn, dev, host_pool, world, rank)
    except Exception as ex:
        if world > 1:
            raise
        train_step = {"error": repr(ex)[:300]}
    scenes = SCENES_PER_GPU * world * args.steps
    value = scenes / (ms * 1e-3)
    e2e_value = scenes / (ms_e2e * 1e-3)

    if rank == 0:
        hbm_peak, tensor_peak, peak_kind = peaks()
        # profile the calls under the SAME options the timed graphs captured (FPS variant, fused-SA grid)
        with L.thread_options(**graphs[0].options):
            ops = per_op_profile(torch, L, model, dev_pool[1])
        step_ms = ms / args.steps
        # dominant kernel: the SA1 furthest-point sampling call (first FPS call of the forward)
        fps_key = next((k for k in ops if k[0].startswith("pn2_furthest_point_sampling") and k[1] == 0), None)
        roofline, roofline_ops = None, []
        if fps_key is not None:
            t_fps = ops[fps_key] * 1e-3
            b = SCENES_PER_GPU
            bytes_alg = b * (12 * N_POINTS + 16 * 2048)  # xyz in, inds + new_xyz out
            updates = b * N_POINTS * 2047                # N (m - 1) distance updates per scene ...
            flops_alg = updates * 10                     # ... at 10 flop each (3 sub, mul, 2 fma, min, cmp)
            fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.5 TFLOP/s: SURVEY 8(d)'s third denominator
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic_r02.json")
            if not os.path.exists(tp):
                tp = os.path.join(ROOT, "profiles", "traffic_r01.json")
            if os.path.exists(tp):  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
                traffic = json.load(open(tp)).get(
                    "fps_bucket_kernel_sa1_bytes_per_launch" if throughput_mode else "fps_kernel_sa1_bytes_per_launch")
            ach = flops_alg / t_fps / 1e12
            # FPS is bound by fp32 issue + the latency of its dependent rounds, not by HBM or the tensor pipe
            # (SURVEY 8(d)): the fraction is algorithmic flops over the fp32 FMA peak of the whole GPU.  The
            # bucketed kernel prunes most of the algorithmic updates, so this is work-equivalent throughput,
            # not a pipe utilisation (ncu: profiles/fps_bucket_r02.txt).
            roofline = {"kernel": ("fps_bucket_kernel" if throughput_mode else "fps_kernel") + " (SA1, 40000->2048)",
                        "bound": "fp32-alu/latency", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                        "frac": ach / fp32_peak, "traffic": traffic,
                        "peak_kind": "148 SM x 128 lanes x 2 flop x 1.965 GHz (SURVEY 8d)",
                        "ms": ops[fps_key], "share_of_serialised_forward": ops[fps_key] / sum(ops.values()),
                        "us_per_sample": ops[fps_key] * 1e3 / 2047,
                        "step_level_frac": flops_alg / (step_ms * 1e-3) / 1e12 / fp32_peak,
                        "hbm_note": {"algorithmic_bytes": bytes_alg, "achieved_gbs": bytes_alg / t_fps / 1e9,
                                     "peak_gbs": hbm_peak, "frac": bytes_alg / t_fps / 1e9 / hbm_peak, "peak_kind": peak_kind},
                        "sms_occupied": SCENES_PER_GPU * (1 if throughput_mode else 16),
                        "note": "timed alone with CUDA events around the C-ABI call (one launch, eager, same options as "
                                "the captured graphs); in the timed region ~6 of these run concurrently with other batches' "
                                "kernels, which is what step_level_frac accounts for"}
        work = algorithmic_work()
        for (name, k), v in sorted(ops.items(), key=lambda kv: -kv[1]):
            row = {"call": name, "nth": k, "ms": round(v, 4), "share_of_serialised_forward": round(v / sum(ops.values()), 4)}
            w = work.get((name, k))
            if w:
                bound, amount = w
                if bound == "tensor":
                    ach = amount / (v * 1e-3) / 1e12
                    row.update({"bound": "tensor", "achieved": round(ach, 3), "unit": "TFLOP/s",
                                "frac": round(ach / tensor_peak, 5)})
                else:
                    ach = amount / (v * 1e-3) / 1e9
                    row.update({"bound": bound, "achieved": round(ach, 2), "unit": "GB/s",
                                "frac": round(ach / hbm_peak, 5)})
            roofline_ops.append(row)
        line = {"metric": "scenes/sec Pointnet2Backbone fwd (40k pts)", "value": value, "unit": "scenes/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": "Pointnet2Backbone fwd, %d scenes/GPU x %d pts, C_in=%d, eval" % (
                    SCENES_PER_GPU, N_POINTS, C_IN), "scenes_per_gpu": SCENES_PER_GPU, "points": N_POINTS,
                    "c_in": C_IN, "l2": "rotating pool of %d input batches (%.0f MB > L2)" % (
                        POOL, POOL * SCENES_PER_GPU * N_POINTS * 4 * (3 + C_IN) / 1e6),
                    "parallelism": "scene-sharded dp%d, no data-path collective" % world,
                    "execution": execution, "batches_in_flight": depth,
                    "e2e_lane_priority_levels": levels, "e2e_host_calls_per_step": 3, "numa": numa,
                    "fps_algorithm": "bucketed single-SM (throughput mode)" if throughput_mode else "16-CTA cluster (latency mode)"},
                "e2e": {"value": e2e_value, "unit": "scenes/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                        "host_issue_ms": round(host_issue_ms_e2e, 3)},
                "host_issue_ms": round(host_issue_ms, 3),  # wall time the host spent enqueueing the K steps (rank 0)
                "timed_region_attempts": {"device_resident": attempts_dev, "e2e": attempts_e2e,
                                          "rule": "a region whose enqueue took > 35 % of its device time (host stall: the "
                                                  "GPU was not kept fed) is measured again, at most twice; the last "
                                                  "attempt is the one reported"},
                "gpu_launches": launches, "clocks": clk, "roofline": roofline, "roofline_ops": roofline_ops[:24],
                "roofline_ops_note": "per C-ABI call, timed alone (eager, CUDA events, median of 5) under the options the "
                                     "timed graphs captured (%s); shares are of the SERIALISED per-op sum, not of the "
                                     "pipelined step" % json.dumps(graphs[0].options)}
        line["grounding_step_forward"] = grounding
        line["train_step"] = train_step
        if world == 1:
            try:
                line["ref_cuda_ext"] = reference_cuda_ext_leg(torch, pkg, model, dev_pool, min(args.steps, 5))
            except Exception as ex:  # the baseline must never break the product's line
                line["ref_cuda_ext"] = {"error": repr(ex)[:200]}
        if world == 1:
            try:  # configs[1] in full: backbone + Hough voting + vote aggregation
                fe = pkg.VoteNetFrontEnd(input_feature_dim=C_IN)
                fe.backbone_net = model
                syn.randomize_bn_stats(fe.vgen); syn.randomize_bn_stats(fe.proposal)
                fe = fe.eval().to(dev)
                with torch.no_grad():
                    for s in range(3):
                        fe({"point_clouds": dev_pool[s]})
                    torch.cuda.synchronize()
                    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    v0.record()
                    for s in range(20):
                        fe({"point_clouds": dev_pool[(3 + s) % POOL]})
                    v1.record()
                    torch.cuda.synchronize()
                vms = v0.elapsed_time(v1) / 20
                line["votenet_front_end"] = {"value": SCENES_PER_GPU / (vms * 1e-3), "unit": "scenes/s", "ms_per_step": vms,
                                             "what": "backbone + VotingModule + L2 norm + vote aggregation (256 clusters), "
                                                     "eager, one batch in flight (host-launch-bound: follows the host's speed "
                                                     "in this process, unlike the graphed figure)"}
                # the same front end the way the backbone is measured: CUDA-graph replays, `depth` batches in flight
                fgraphs = [pkg.GraphedBackbone(fe, dev_pool[0], throughput=throughput_mode) for k in range(depth)]
                if all(g.graph is not None for g in fgraphs):
                    def fe_steps(count):
                        for st in lanes:
                            st.wait_stream(main_stream)
                        for s in range(count):
                            with torch.cuda.stream(lanes[s % depth]):
                                fgraphs[s % depth](dev_pool[s % POOL])
                        for st in lanes:
                            main_stream.wait_stream(st)
                    settle(torch, fe_steps)
                    v0.record()
                    fe_steps(200)
                    v1.record()
                    torch.cuda.synchronize()
                    gms = v0.elapsed_time(v1) / 200
                    line["votenet_front_end"].update({"graphed_value": SCENES_PER_GPU / (gms * 1e-3),
                                                      "graphed_ms_per_step": gms, "graphed_batches_in_flight": depth})
                del fgraphs
            except Exception as ex:
                line["votenet_front_end"] = {"error": repr(ex)[:200]}
            try:  # configs[1] to the end of the detection branch: + ROI head + box decode (ProposalModule)
                import numpy as np
                det = pkg.VoteNetDetector(input_feature_dim=C_IN, mean_size_arr=np.ones((18, 3), dtype=np.float32))
                det.backbone_net = model
                det.vgen = fe.vgen
                det.proposal.vote_aggregation = fe.proposal.vote_aggregation
                syn.randomize_bn_stats(det.proposal.proposal)
                det = det.eval().to(dev)
                dgraphs = [pkg.GraphedBackbone(det, dev_pool[0], throughput=throughput_mode) for k in range(depth)]
                if all(g.graph is not None for g in dgraphs):
                    def det_steps(count):
                        for st in lanes:
                            st.wait_stream(main_stream)
                        for s in range(count):
                            with torch.cuda.stream(lanes[s % depth]):
                                dgraphs[s % depth](dev_pool[s % POOL])
                        for st in lanes:
                            main_stream.wait_stream(st)
                    settle(torch, det_steps)
                    v0.record()
                    det_steps(200)
                    v1.record()
                    torch.cuda.synchronize()
                    dms = v0.elapsed_time(v1) / 200
                    line["votenet_detector"] = {"value": SCENES_PER_GPU / (dms * 1e-3), "unit": "scenes/s",
                                                "ms_per_step": dms, "batches_in_flight": depth,
                                                "what": "backbone + VotingModule + ProposalModule (vote aggregation, ROI head, "
                                                        "box decode to 256 x 8 corners), CUDA-graph replays"}
                del dgraphs
            except Exception as ex:
                line["votenet_detector"] = {"error": repr(ex)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            try:  # achieved error of the captured graph at the bench configuration, one scene against the oracle
                from oracle import oracle as O
                sd0 = {k: v.cpu() for k, v in model.state_dict().items()}
                got = graphs[0](dev_pool[0])
                torch.cuda.synchronize()
                want = O.backbone(host_pool[0, :1].clone(), sd0)
                num = {"scene": "batch 0, scene 0 (checker: oracle.backbone, fp32 reference semantics)"}
                for k in ("sa1_inds", "sa2_inds", "fp2_inds", "sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz"):
                    num[k + "_bit_exact"] = bool(torch.equal(got[k][:1].cpu(), want[k]))
                for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
                    a, w = got[k][:1].float().cpu(), want[k]
                    num[k + "_max_norm_err"] = float((a - w).abs().max() / w.abs().max())
                    num[k + "_rel_rms_err"] = float((a - w).pow(2).mean().sqrt() / w.pow(2).mean().sqrt())
                line["numerics"] = num
            except Exception as ex:
                line["numerics"] = {"error": repr(ex)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            O.ext.threads = cores
            sd = {k: v.cpu() for k, v in model.state_dict().items()}
            one = host_pool[0, :1].clone()
            with torch.no_grad():
                O.backbone(one, sd)  # warm-up
                t0 = time.perf_counter()
                n_cpu = 0
                while n_cpu < 5 and time.perf_counter() - t0 < 20.0:
                    O.backbone(host_pool[0, n_cpu % SCENES_PER_GPU:n_cpu % SCENES_PER_GPU + 1].clone(), sd)
                    n_cpu += 1
                dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n_cpu / dt, "unit": "scenes/s", "cores": cores, "kind": "port",
                                    "sample": "%d scenes, B=1, N=40000, C_in=1, fp32 (oracle C ops + torch layers)" % n_cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
