"""oracle/ -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package; the product (closerlook3d_b200/) never does.

  cl3d_oracle.c   plain-C restatement of the reference's five CUDA ops (cites file:line)
  ext.py          the C restatement behind the reference's pybind surface (`pt_custom_ops._ext`)
  la_oracle.py    torch-CPU restatement of pt_utils.py + local_aggregation_operators.py (cites file:line)
  ref_loader.py   imports the UNMODIFIED reference python modules from /root/reference (container only)
  build_ref.py    compiles the UNMODIFIED reference CUDA extension into oracle/_ref (GPU-side pin)
  make_golden.py  generates tests/golden/*.pt from the reference python modules
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "cl3d_oracle.c")
_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off -fopenmp cl3d_oracle.c -> oracle/liboracle.so"""
    if (not force) and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared", "-o", _SO, _SRC,
           "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib
