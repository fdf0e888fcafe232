"""Developer tool: build the library with -DPN2_FB_DEBUG (per-phase cycle counters printed by the
bucketed FPS kernel) into tools/_build and run SA1-sized FPS once."""
import ctypes, importlib, os, subprocess, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "3dvlp_b200", "csrc")
CPW = os.environ.get("PN2_FB_CPW", "")  # candidate groups per warp to try (1 / 2 / 4); default = the library's
OUT = os.path.join(ROOT, "tools", "_build", "libfb_debug%s.so" % CPW)
def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("fps.cu", "fps_bucket.cu", "fps_prefix.cu", "abi.cu")]
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-DPN2_FB_DEBUG"] + (["-DPN2_FB_CPW=" + CPW] if CPW else []) + ["-gencode",
                           "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", OUT] + srcs +
                          ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
if not torch.cuda.is_available():
    build(); print("built"); sys.exit(0)
sys.path.insert(0, ROOT)
syn = importlib.import_module("3dvlp_b200.synthetic")
os.environ.setdefault("PN2_FPS_BUCKET", "1")  # the bucketed kernel is what this tool times
lib = ctypes.CDLL(OUT)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
x = syn.make_batch(0, 8, n, 0).cuda().contiguous()
idx = torch.zeros(8, m, dtype=torch.int32, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    lib.pn2_furthest_point_sampling(8, n, m, ctypes.c_void_p(x.data_ptr()), None, ctypes.c_void_p(idx.data_ptr()), st)
torch.cuda.synchronize()
