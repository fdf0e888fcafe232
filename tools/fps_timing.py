"""Developer tool: per-phase cycle breakdown of the FPS round loop.

Builds fps.cu + abi.cu with -DPN2_FPS_TIMING into tools/_build/libfps_timing.so (clock64() at the
phase boundaries of CTA 0 / thread 0) and prints average cycles per round for each phase:
  A update+thread argmax | B warp redux+record | sync | C+D cta reduce+st.async | wait mbarrier | E reduce+decode
Usage (GPU box):  python tools/fps_timing.py [n] [m] [b]      env PN2_FPS_CLUSTER / PN2_FPS_THREADS apply.
"""
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "3dvlp_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_build", "libfps_timing.so")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("fps.cu", "abi.cu")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) > os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-DPN2_FPS_TIMING", "-gencode",
                           "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", OUT] + srcs +
                          ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])


def main():
    build()
    if not torch.cuda.is_available():
        print("built", OUT)
        return
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    b = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    lib = ctypes.CDLL(OUT)
    if os.environ.get("PN2_OCC"):
        pass
    if os.environ.get("PN2_SCENES"):
        import importlib
        sys.path.insert(0, ROOT)
        syn = importlib.import_module("3dvlp_b200.synthetic")
        x = syn.make_batch(0, b, n, 0).cuda().contiguous()
    else:
        x = (torch.rand(b, n, 3, device="cuda") * 6 - 3).contiguous()
    idx = torch.zeros(b, m, dtype=torch.int32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        rc = lib.pn2_furthest_point_sampling(b, n, m, ctypes.c_void_p(x.data_ptr()), None, ctypes.c_void_p(idx.data_ptr()), st)
        assert rc == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.pn2_furthest_point_sampling(b, n, m, ctypes.c_void_p(x.data_ptr()), None, ctypes.c_void_p(idx.data_ptr()), st)
    e1.record()
    torch.cuda.synchronize()
    out = (ctypes.c_longlong * 7)()
    assert lib.pn2_fps_timing(out) == 0
    names = ["A update", "B warp-redux", "sync", "C+D cta+send", "wait", "E accept+sync"]
    ph = list(out)[:6]
    rounds = max(1, out[6])
    print("n=%d m=%d b=%d cluster=%s threads=%s  %.3f ms  rounds=%d (%.2f samples/round)  %.3f us/round" % (
        n, m, b, os.environ.get("PN2_FPS_CLUSTER", "auto"), os.environ.get("PN2_FPS_THREADS", "auto"),
        e0.elapsed_time(e1), rounds, (m - 1) / rounds, e0.elapsed_time(e1) * 1e3 / rounds))
    print("  cycles/round: " + " | ".join("%s %.0f" % (nm, v / rounds) for nm, v in zip(names, ph)) +
          " | total %.0f" % (sum(ph) / rounds))


if __name__ == "__main__":
    main()
