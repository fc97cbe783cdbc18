"""TEST INFRASTRUCTURE: compile the reference's own native kernel.

Cythonises ``/root/reference/sparray/_merge.pyx`` *where it lies* (the
reference's only native file) and links it into
``oracle/_ref/_merge*.so``.  Nothing is copied into the repo: the generated C
file and the shared object both land in ``oracle/_ref/`` (git-ignored, but
shipped to the GPU box by gpurun).  We do not run the reference's setup.py.

The module is importable as top-level ``_merge`` -- which is exactly what the
reference's ``compat.py:106`` tries to import.

Usage: python oracle/build_ref.py [reference_root]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def so_path():
    return os.path.join(OUT, "_merge" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(ref_root="/root/reference", force=False):
    """Returns the path of the built module, or None if the reference tree
    (or Cython) is not available here."""
    pyx = os.path.join(ref_root, "sparray", "_merge.pyx")
    target = so_path()
    if os.path.exists(target) and not force:
        return target
    if not os.path.exists(pyx):
        return None
    try:
        import Cython  # noqa: F401
        import numpy as np
    except ImportError:
        return None
    os.makedirs(OUT, exist_ok=True)
    c_file = os.path.join(OUT, "_merge.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file])
    cc = os.environ.get("CC", "gcc")
    cmd = [cc, "-O2", "-fPIC", "-shared", "-w",
           "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(),
           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
           c_file, "-o", target]
    subprocess.check_call(cmd)
    os.remove(c_file)  # keep only the binary: no reference-derived source stays
    return target


def load():
    """Import the built reference kernel module (or None if it was not built)."""
    target = so_path()
    if not os.path.exists(target):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_merge", target)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    print(build(root, force=True))
