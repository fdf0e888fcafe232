"""Build recipe for ``oracle/_ref`` -- TEST INFRASTRUCTURE, never shipped or imported by the product.

Compiles the *unmodified* reference CUDA extension (the nine ``pointnet2._ext`` ops,
/root/reference/lib/pointnet2/_ext_src, build recipe setup.py:19-34) from the sources where
they lie, for sm_100a, into ``oracle/_ref/pointnet2_ref_ext.so``.  Only the binary lands in the
repo tree (``oracle/_ref/`` is git-ignored but travels to the GPU box with gpurun); no reference
source is copied.  On the GPU box the .so is the strongest pin for the CPU oracle: the parity
tests run the reference's own kernels on the same inputs (tests/test_ref_ext_gpu.py).

``install_reference_stack`` additionally lays the reference's own Python layers of the hot path
(lib/pointnet2/{pointnet2_utils,pointnet2_modules,pytorch_utils}.py, models/base_module/backbone_module.py,
models/base_module/voting_module.py, utils/nn_distance.py -- the list of SURVEY.md 8c) out under
``baseline/_ref/`` with their package-relative paths: the reference has no setup.py at its root, so this
file copy IS its install (``pip install /root/reference`` has nothing to build).  ``baseline/_ref/`` is
git-ignored -- no reference source enters the history -- but not gpurun-ignored: bench.py's GPU baseline
(``ref_cuda_ext``) imports the reference's stock ``Pointnet2Backbone`` from there and runs it on the
reference's own compiled ``_ext``.  Never imported by the product.

Run in the build container only (needs /root/reference):  python oracle/build_ref.py
"""
import glob
import os
import sys

REF_SRC = "/root/reference/lib/pointnet2/_ext_src"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
NAME = "pointnet2_ref_ext"


def build(verbose=False):
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: only the prebuilt file is used
    os.makedirs(OUT_DIR, exist_ok=True)
    so = os.path.join(OUT_DIR, NAME + ".so")
    srcs = sorted(glob.glob(REF_SRC + "/src/*.cpp") + glob.glob(REF_SRC + "/src/*.cu"))
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=srcs, extra_include_paths=[REF_SRC + "/include"],
         extra_cflags=["-O2"],
         extra_cuda_cflags=["-O2", "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=OUT_DIR, verbose=verbose, is_python_module=True)
    # keep only the binary (ninja files / objects are scratch)
    for f in os.listdir(OUT_DIR):
        if not f.endswith(".so"):
            try:
                os.remove(os.path.join(OUT_DIR, f))
            except OSError:
                pass
    return so


REF_ROOT = "/root/reference"
BASELINE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
STACK = ("lib/pointnet2/pointnet2_utils.py", "lib/pointnet2/pointnet2_modules.py", "lib/pointnet2/pytorch_utils.py",
         "models/base_module/backbone_module.py", "models/base_module/voting_module.py", "utils/nn_distance.py")


def install_reference_stack():
    """Copy the reference's hot-path Python layers, unmodified, into baseline/_ref/ (git-ignored)."""
    if not os.path.isdir(REF_ROOT):
        return None  # GPU box: the files travelled with the snapshot
    import shutil
    for rel in STACK:
        dst = os.path.join(BASELINE_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_ROOT, rel), dst)
    return BASELINE_DIR


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
    print(install_reference_stack())
