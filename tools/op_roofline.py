"""Per-op roofline table of the nine `_ext` operators + nn_distance, each timed ALONE through the
C ABI at sizes where one launch is long enough to be measured (the backbone's own SA3/SA4/FP sizes
finish inside launch latency).  Algorithmic bytes / flops follow SURVEY.md 8(d); peaks come from
MEASURED_PEAKS.json (burst figures: each kernel is timed alone).  Inputs are rotated through a pool
larger than L2 so no launch re-reads a warm input.

Usage: python tools/op_roofline.py [out.json]
"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
ext = importlib.import_module("3dvlp_b200._ext")
nnd = importlib.import_module("3dvlp_b200.nn_distance")
syn = importlib.import_module("3dvlp_b200.synthetic")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm, fp32 = 6539.9, 148 * 128 * 2 * 1.965e9 / 1e12
    if os.path.exists(p):
        d = json.load(open(p))
        hbm = float(d.get("hbm_gbs", hbm))
    return hbm, fp32


def timed(fn, pool, reps):
    for i in range(3):
        fn(pool[i % len(pool)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(pool[i % len(pool)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    hbm, fp32 = peaks()
    rows = []

    def add(name, shape, ms, nbytes, flops, bound):
        gbs = nbytes / ms / 1e6
        tfl = flops / ms / 1e9
        r = {"op": name, "shape": shape, "ms": round(ms, 4), "bound": bound, "algorithmic_MB": round(nbytes / 1e6, 2),
             "GBps": round(gbs, 1), "hbm_frac": round(gbs / hbm, 4), "fp32_TFLOPs": round(tfl, 2),
             "fp32_frac": round(tfl / fp32, 4)}
        rows.append(r)
        print(json.dumps(r), flush=True)

    g = torch.Generator().manual_seed(0)
    # ---- calibration: what this box gives a pure write and a copy of 1 GB ---------------------
    big = torch.empty(256 << 20, device=dev)
    src = torch.rand(256 << 20, device=dev)
    add("calibration: fill_ (pure write)", "1 GiB", timed(lambda x: big.fill_(1.0), [0], 10), big.numel() * 4, 0, "hbm")
    add("calibration: copy_ (read + write)", "1 GiB + 1 GiB", timed(lambda x: big.copy_(src), [0], 10), big.numel() * 8, 0, "hbm")
    del big, src
    # ---- scenes: the bench generator at three sizes -----------------------------------------
    for B, N, npoint in ((8, 40000, 2048), (8, 100000, 4096), (32, 20000, 2048)):
        pool = [syn.make_batch(100 * k, B, N, 1)[..., :3].contiguous().to(dev) for k in range(3)]
        shape = "B=%d N=%d m=%d" % (B, N, npoint)
        # FPS: 10 flop per point-update, bytes 12N + 4m
        ms = timed(lambda x: ext.furthest_point_sampling(x, npoint), pool, 5)
        add("furthest_point_sampling", shape, ms, B * (12 * N + 4 * npoint), 10.0 * B * N * (npoint - 1), "latency/fp32")
        inds = [ext.furthest_point_sampling(x, npoint) for x in pool]
        new_xyz = [torch.gather(x, 1, i.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous() for x, i in zip(pool, inds)]
        pairs = list(zip(new_xyz, pool))
        for radius, ns in ((0.2, 64), (0.4, 32)):
            ms = timed(lambda p: ext.ball_query(p[0], p[1], radius, ns), pairs, 10)
            add("ball_query r=%.1f ns=%d" % (radius, ns), shape, ms, B * (12 * N + 12 * npoint + 4 * npoint * ns),
                9.0 * B * N * npoint, "fp32 (brute-force test count)")
        # three_nn: unknown = all points of a 20k subset, known = the samples
        unk = [x[:, :20000].contiguous() for x in pool]
        pairs3 = list(zip(unk, new_xyz))
        ms = timed(lambda p: ext.three_nn(p[0], p[1]), pairs3, 10)
        add("three_nn", "B=%d n=20000 m=%d" % (B, npoint), ms, B * (12 * (20000 + npoint) + 24 * 20000),
            9.0 * B * 20000 * npoint, "fp32")

    # ---- feature movers ---------------------------------------------------------------------
    for B, C, N, M, S in ((8, 128, 40000, 2048, 64), (8, 256, 2048, 1024, 32), (8, 3, 40000, 2048, 64)):
        feats = [torch.rand(B, C, N, generator=g).to(dev) for _ in range(max(2, int(400e6 / (B * C * N * 4)) + 1))][:6]
        xyz = syn.make_batch(7, B, max(N, 4096), 1)[:, :N, :3].contiguous().to(dev)
        ind = ext.furthest_point_sampling(xyz, M)
        nx = torch.gather(xyz, 1, ind.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        idx = ext.ball_query(nx, xyz, 0.2 if N > 4096 else 0.4, S)
        shape = "B=%d C=%d N=%d M=%d S=%d" % (B, C, N, M, S)
        ms = timed(lambda f: ext.gather_points(f, ind), feats, 20)
        add("gather_points", shape, ms, B * (2 * C * M * 4 + 4 * M), 0, "hbm")
        ms = timed(lambda f: ext.group_points(f, idx), feats, 10)
        # unique rows read once (they stay in L2 across a ball) + the grouped write
        add("group_points", shape, ms, B * (C * N * 4 + C * M * S * 4 + 4 * M * S), 0, "hbm")
        go = [torch.rand(B, C, M, S, generator=g).to(dev) for _ in range(2)]
        ms = timed(lambda f: ext.group_points_grad(f, idx, N), go, 5)
        add("group_points_grad", shape, ms, B * (C * N * 4 + C * M * S * 4 + 4 * M * S), 0, "hbm (atomics)")
        go2 = [torch.rand(B, C, M, generator=g).to(dev) for _ in range(2)]
        ms = timed(lambda f: ext.gather_points_grad(f, ind, N), go2, 10)
        add("gather_points_grad", shape, ms, B * (C * N * 4 + C * M * 4 + 4 * M), 0, "hbm (zero-fill + atomics)")
        del feats, go, go2

    for B, C, m, n in ((8, 256, 256, 512), (8, 256, 512, 1024), (32, 256, 2048, 40000)):
        known = [torch.rand(B, C, m, generator=g).to(dev) for _ in range(4)]
        kx = torch.rand(B, m, 3, generator=g).to(dev)
        ux = torch.rand(B, n, 3, generator=g).to(dev)
        d2, idx = ext.three_nn(ux, kx)
        w = 1.0 / (torch.sqrt(d2) + 1e-8)
        w = (w / w.sum(2, keepdim=True)).contiguous()
        shape = "B=%d C=%d m=%d n=%d" % (B, C, m, n)
        ms = timed(lambda f: ext.three_interpolate(f, idx, w), known, 20)
        add("three_interpolate", shape, ms, B * (4 * C * m + 24 * n + 4 * C * n), 5.0 * B * C * n, "hbm")
        go = [torch.rand(B, C, n, generator=g).to(dev) for _ in range(2)]
        ms = timed(lambda f: ext.three_interpolate_grad(f, idx, w, m), go, 10)
        add("three_interpolate_grad", shape, ms, B * (4 * C * m + 24 * n + 4 * C * n), 0, "hbm (atomics)")
        del known, go

    for B, N, M, C, kw in ((8, 256, 256, 3, {}), (8192, 1, 3, 3, {"l1": True}), (64, 256, 1, 3, {}), (8, 2048, 2048, 3, {}),
                           (8, 20000, 20000, 3, {})):
        a = [torch.rand(B, N, C, generator=g).to(dev) for _ in range(2)]
        b = torch.rand(B, M, C, generator=g).to(dev)
        ms = timed(lambda x: nnd.nn_distance(x, b, **kw), a, 10)
        add("nn_distance" + (" l1" if kw else ""), "B=%d N=%d M=%d C=%d" % (B, N, M, C), ms,
            B * (4 * C * (N + M) + 12 * (N + M)), 3.0 * C * B * N * M, "hbm / fp32")

    out = {"device": torch.cuda.get_device_name(0), "hbm_peak_GBps": hbm, "fp32_peak_TFLOPs": round(fp32, 2), "rows": rows}
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
