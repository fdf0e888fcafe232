"""Writes tests/golden/reference_ext.npz: outputs of the UNMODIFIED reference CUDA extension
(oracle/_ref/pointnet2_ref_ext.so = /root/reference/lib/pointnet2/_ext_src compiled for sm_100a by
oracle/build_ref.py) on the seeded inputs of ref_ext_cases.py.  Needs a GPU:

    gpurun -- 'python tests/golden/make_golden_ref_ext.py gpurun_out/reference_ext.npz'

then copy gpurun_out/reference_ext.npz to tests/golden/.  With `--oracle` the CPU oracle runs instead
(a dry run of the script; never used to produce the committed fixture)."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_ext_cases as cases  # noqa: E402

dst = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else os.path.join(HERE, "reference_ext.npz")
if "--oracle" in sys.argv:
    from oracle import oracle as O
    O.build()
    ext, dev, producer = O.ext, torch.device("cpu"), "oracle (dry run)"
else:
    so = os.path.join(ROOT, "oracle", "_ref", "pointnet2_ref_ext.so")
    spec = importlib.util.spec_from_file_location("pointnet2_ref_ext", so)
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    dev = torch.device("cuda:0")
    producer = "reference _ext (sm_100a build) on " + torch.cuda.get_device_name(0)
out = cases.run_all(ext, dev)
out["_producer"] = np.array(producer)
np.savez_compressed(dst, **out)
print("wrote %s: %d arrays, %d bytes, producer: %s" % (dst, len(out), os.path.getsize(dst), producer))
