"""Load the UNMODIFIED reference package from /root/reference for golden-vector
generation (build container only; /root/reference does not exist on the GPU box).

The reference predates numpy 2: importing works, but the ctor
fails on ``np.array(..., copy=False)`` and a few removed aliases.  The six
mechanical numpy-2 substitutions listed in SURVEY.md section 8(c) are applied to
the source text IN MEMORY before exec -- nothing from the reference is written
to disk.  Behaviour on numpy 1.x is unchanged by them.

``load_reference(use_cython=True)`` additionally makes the reference's own
Cython kernel (built by oracle/build_ref.py into oracle/_ref/) importable as the
top-level module ``_merge``, which is the name ``compat.py:106`` looks for
(SURVEY.md F3); with ``use_cython=False`` the numpy fallbacks are selected.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("SPARRAY_REFERENCE_ROOT", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

_SUBS = {
    "flat.py": [
        ("np.array(indices, dtype=int, copy=False)", "np.asarray(indices, dtype=int)"),
        ("np.array(data, copy=False)", "np.asarray(data)"),
        ("np.array(arr, copy=False)", "np.asarray(arr)"),
        ("np.array(idx, copy=False, subok=True, order='A')", "np.asanyarray(idx)"),
        ("np.product(", "np.prod("),
    ],
    "base.py": [
        ("np.float_", "np.float64"),
        ("np.complex_", "np.complex128"),
    ],
    "compat.py": [
        ("np.array(array, copy=False, subok=subok)", "np.array(array, copy=None, subok=subok)"),
    ],
}


def available():
    return os.path.exists(os.path.join(REF_ROOT, "sparray", "flat.py"))


def _exec_module(pkg, name, fname):
    path = os.path.join(REF_ROOT, "sparray", fname)
    with open(path) as fh:
        src = fh.read()
    for old, new in _SUBS.get(fname, []):
        src = src.replace(old, new)
    mod = types.ModuleType(pkg.__name__ + "." + name)
    mod.__package__ = pkg.__name__
    mod.__file__ = path
    sys.modules[mod.__name__] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    setattr(pkg, name, mod)
    return mod


def load_reference(use_cython=True, alias="sparray_ref"):
    """Returns the reference package as a module object exposing FlatSparray,
    is_sparray, compat, flat, base."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    alias = alias + ("_cy" if use_cython else "_np")
    if alias in sys.modules:
        return sys.modules[alias]
    saved = sys.modules.pop("_merge", None)
    blocker = None
    if use_cython:
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import build_ref
        sys.path.pop(0)
        if build_ref.build(REF_ROOT) is None:
            raise RuntimeError("could not build the reference Cython kernel")
        sys.modules["_merge"] = build_ref.load()
    else:
        # make `from _merge import ...` and the pyximport retry both fail, so
        # compat.py:117-120 picks the pure-numpy fallbacks
        class _Block:
            def find_spec(self, fullname, path=None, target=None):
                if fullname in ("_merge", "pyximport"):
                    raise ImportError("blocked for the numpy-fallback oracle")
                return None
        blocker = _Block()
        sys.meta_path.insert(0, blocker)
    try:
        pkg = types.ModuleType(alias)
        pkg.__path__ = []
        pkg.__package__ = alias
        sys.modules[alias] = pkg
        compat = _exec_module(pkg, "compat", "compat.py")
        base = _exec_module(pkg, "base", "base.py")
        flat = _exec_module(pkg, "flat", "flat.py")
        pkg.FlatSparray = flat.FlatSparray
        pkg.is_sparray = base.is_sparray
        want = "cython_function_or_method" if use_cython else "function"
        got = type(compat.union1d_sorted).__name__
        assert got == want, "wrong reference kernel path selected: %s" % got
    finally:
        if blocker is not None:
            sys.meta_path.remove(blocker)
        sys.modules.pop("_merge", None)
        if saved is not None:
            sys.modules["_merge"] = saved
    return pkg
