"""Developer tool: build the library with -DPN2_SA_TIMING (per-phase cycle counters printed by two CTAs of every
fused SA launch) into tools/_build and run one backbone forward.  Usage: python tools/sa_debug.py [throughput]"""
import importlib, os, subprocess, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "3dvlp_b200", "csrc")
EXP = os.environ.get("PN2_SA_EXP", "")  # e.g. NOLD / NOMAX / NORED: -DPN2_SA_EXP_<name>
OUT = os.path.join(ROOT, "tools", "_build", "libsa_debug%s.so" % EXP)


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-DPN2_SA_TIMING"] + (["-DPN2_SA_EXP_" + EXP] if EXP else []) + ["-gencode",
                           "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", OUT] + srcs +
                          ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])


if not torch.cuda.is_available():
    build(); print("built"); sys.exit(0)
os.environ["PN2_LIB_PATH"] = OUT
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
dev = torch.device("cuda:0")
torch.manual_seed(42)
bb = pkg.Pointnet2Backbone(input_feature_dim=1)
syn.randomize_bn_stats(bb)
bb = bb.eval().to(dev)
pc = syn.make_batch(0, 8, 40000, 1).to(dev)
opts = dict(fps_bucket=1, sa_tiles_per_group=16) if len(sys.argv) > 1 and sys.argv[1] == "throughput" else {}
with torch.no_grad(), pkg.thread_options(**opts):
    for i in range(2):
        print("---- forward", i, flush=True)
        bb({"point_clouds": pc})
        torch.cuda.synchronize()
