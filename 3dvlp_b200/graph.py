"""CUDA-graph replay of the eval-mode backbone (GraphedBackbone) and of a whole training step
(GraphedTrainStep).

A backbone forward is ~40 short kernel launches on two streams; at 2 ms per step the Python / driver
launch path is of the same order as the GPU work.  Shapes are static (B scenes x N points), every
kernel is stream-ordered (the ball-query workspace comes from the stream-ordered allocator), so the
whole forward -- including the fork/join of the sampling stream -- is captured ONCE into a CUDA graph
and replayed per batch: one launch per step, no tracing compiler involved.
"""
import torch


class GraphedBackbone:
    """``g = GraphedBackbone(backbone, example_batch); out = g(point_clouds)``.

    `example_batch` (B, N, 3+C) fp32 CUDA fixes the shapes.  ``g(pc)`` copies `pc` (device or pinned
    host tensor) into the static input, replays the graph on the current stream and returns the dict
    of STATIC output tensors (same keys as Pointnet2Backbone.forward); they are overwritten by the
    next call.  Falls back to eager execution (``self.graph is None``) if capture is not possible.
    """

    def __init__(self, backbone, example_batch, warmup=3, throughput=False, priority=0, pack=None):
        if backbone.training:
            raise RuntimeError("GraphedBackbone needs the backbone in eval() mode")
        if not example_batch.is_cuda or example_batch.dtype != torch.float32:
            raise RuntimeError("CPU not supported")
        self.backbone = backbone
        # throughput=True captures the single-SM bucketed FPS (meant for several graphs in flight on
        # separate streams); the default captures the low-latency 16-SM cluster FPS
        # The options are set as PER-THREAD overrides (pn2_set_option_thread): a concurrent eager call on
        # another host thread keeps launching what the process-wide options say.
        from . import _lib
        # throughput=True also launches the fused SA bodies on fewer persistent CTAs (>= 16 tiles per
        # group): their per-CTA prologue is SM-time that other batches in flight could use
        opts = dict(fps_bucket=1, sa_tiles_per_group=_lib.get_option("sa_tiles_per_group") or 16) if throughput \
            else dict(fps_bucket=0, sa_tiles_per_group=0)
        self.options = opts
        with _lib.thread_options(**opts):
            self.static_in = example_batch.clone().contiguous()
            self.graph = None
            self.static_out = None
            self.capture_error = None
            self.packed, self.pack_layout = None, None
            cur = torch.cuda.current_stream(self.static_in.device)
            # `priority` (0 = default, negative = higher): every kernel of the graph is captured on streams
            # of that priority and keeps it on replay.  A serving loop with several graphs in flight gives
            # them descending priorities so that batches complete one after the other -- their device->host
            # reads overlap the compute of the batches behind them -- instead of all at once.
            s = torch.cuda.Stream(device=self.static_in.device, priority=priority)
            self.priority = s.priority
            s.wait_stream(cur)
            with torch.no_grad(), torch.cuda.stream(s):
                for _ in range(max(1, warmup)):  # fills every host-side cache (plans, occupancy, pools)
                    backbone({"point_clouds": self.static_in})
            cur.wait_stream(s)
            torch.cuda.synchronize(self.static_in.device)
            try:
                g = torch.cuda.CUDAGraph()
                inputs = {"point_clouds": self.static_in}
                if pack:
                    # `pack` = the backbone outputs a serving loop reads back every step (fp2_features, fp2_xyz,
                    # fp2_inds): they live in ONE contiguous byte buffer, so the host issues one device->host
                    # copy per step instead of one per tensor.  The final linear kernel writes fp2_features
                    # straight into its slice; the two small tensors are placed by elementwise kernels (a
                    # device-to-device memcpy node would queue behind the PCIe traffic on the copy engines).
                    with torch.no_grad():
                        probe = backbone({"point_clouds": self.static_in})
                    self.pack_layout, off = [], 0
                    for k in pack:
                        nbytes = probe[k].numel() * probe[k].element_size()
                        self.pack_layout.append((k, off, nbytes, tuple(probe[k].shape), probe[k].dtype))
                        off += (nbytes + 255) // 256 * 256
                    del probe
                    torch.cuda.synchronize(self.static_in.device)
                    self.packed = torch.empty(off, device=self.static_in.device, dtype=torch.uint8)
                    views = self.packed_views(self.packed)
                    if "fp2_features" in views:
                        inputs["_fp2_features_out"] = views["fp2_features"]
                with torch.no_grad(), torch.cuda.graph(g, stream=s):
                    out = backbone(inputs)
                    if pack:
                        for k, v in views.items():
                            if out[k].data_ptr() == v.data_ptr():
                                continue  # written in place by the producing kernel
                            if v.dtype.is_floating_point:
                                torch.mul(out[k], 1.0, out=v)
                            else:
                                torch.bitwise_or(out[k], 0, out=v)
                self.graph, self.static_out = g, out
            except Exception as ex:  # pragma: no cover - depends on driver behaviour
                self.capture_error = repr(ex)
                torch.cuda.synchronize(self.static_in.device)

    def packed_views(self, buf):
        """The tensors of `pack` as views into `buf`, a byte buffer laid out like self.packed (e.g. the pinned
        host buffer the serving loop copied it into)."""
        return {k: buf[o:o + nbytes].view(dt).view(shape) for k, o, nbytes, shape, dt in self.pack_layout}

    def check(self):
        """Synchronise and raise Pn2Error if a kernel of this device bailed out of a bounded wait since the
        last poll (pn2_poll_device_error: 1 = FPS cluster exchange, 2 = tensor-core pipeline).  A bail-out
        leaves partial outputs behind, so serving loops call this at their natural sync points."""
        from . import _lib, tensorcore
        flag = tensorcore.poll_device_error(self.static_in.device)
        if flag:
            raise _lib.Pn2Error("device-side protocol time-out (flag %d): outputs of the last replays are invalid" % flag)

    def __call__(self, point_clouds):
        self.static_in.copy_(point_clouds, non_blocking=True)
        if self.graph is None:
            from . import _lib
            with torch.no_grad(), _lib.thread_options(**self.options):
                return self.backbone({"point_clouds": self.static_in})
        self.graph.replay()
        return self.static_out


class GraphedTrainStep:
    """A whole training iteration -- forward in train() mode, losses, backward, optimizer step -- captured ONCE
    as a CUDA graph and replayed with one launch per iteration.

    The reference's training loop (``lib/joint/solver_3dvlp.py:580-584``: forward, loss, ``backward()``,
    ``optimizer.step()``) issues ~3 900 kernels per iteration at this path's sizes; on a B200 their device time
    is shorter than the time the host needs to launch them.  Every kernel of this library is stream-ordered
    (workspaces come from the stream-ordered allocator, the loss glue has no host synchronisation), so the
    iteration is capturable as it is.

    ``loss_fn()`` runs forward + losses on STATIC input tensors it closes over and returns the scalar loss;
    ``optimizer`` must be capturable (``torch.optim.Adam(..., capturable=True)``).  ``step()`` replays the graph
    and returns the static loss tensor (overwritten by the next replay); new data is fed by ``copy_``-ing it
    into the static inputs first.  BatchNorm running statistics, parameters and optimizer state are updated in
    place by the replay exactly as by an eager iteration.
    """

    def __init__(self, loss_fn, optimizer, warmup=3):
        for grp in optimizer.param_groups:
            if "capturable" in grp and not grp["capturable"]:
                raise RuntimeError("GraphedTrainStep needs a capturable optimizer (e.g. Adam(..., capturable=True))")
        self.loss_fn, self.optimizer = loss_fn, optimizer
        cur = torch.cuda.current_stream()
        s = torch.cuda.Stream()
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):  # host-side caches, lazily created optimizer state, allocator pools
                self._eager()
        cur.wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)  # gradients are (re)created inside the graph's private pool
        with torch.cuda.graph(self.graph):
            loss = loss_fn()
            loss.backward()
            optimizer.step()
            self.loss = loss.detach()

    def _eager(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_fn()
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def step(self):
        self.graph.replay()
        return self.loss
