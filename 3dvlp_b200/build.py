"""Compile the CUDA sources under 3dvlp_b200/csrc into 3dvlp_b200/lib3dvlp_b200.so for sm_100a.

Plain nvcc, no torch headers: the library exposes the C ABI of include/pointnet2_b200.h and is
loaded with ctypes.  It is built IN-TREE so it travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib3dvlp_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--use_fast_math=false"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]  # fast math is never enabled


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "pointnet2_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building lib3dvlp_b200.so")
    link = [NVCC, "-shared", "-o", OUT] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.check_call(link)
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
