"""Generate tests/golden/golden_v1.npz by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every case is evaluated twice -- once on the reference's Cython kernel path
(oracle/_ref, forced on as SURVEY.md F3 describes) and once on its numpy
fallback path -- and the two must agree bit-for-bit before the case is written.
Inputs are small and seeded so the fixture stays well under 1 MB.

Fixture layout: one npz.  ``manifest`` is a JSON list of cases
``{"name", "op", "params", "in": [keys], "out": [keys]}``; arrays are stored
under ``<name>/in/<key>`` and ``<name>/out/<key>``.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

OUT = os.path.join(HERE, "golden_v1.npz")


def rand_sparse(rng, shape, density, dtype):
    size = int(np.prod(shape))
    n = int(round(density * size))
    idx = np.unique(rng.integers(0, max(size, 1), size=n, dtype=np.int64)) if size else np.zeros(0, np.int64)
    if np.issubdtype(np.dtype(dtype), np.integer):
        val = rng.integers(-9, 10, size=len(idx)).astype(dtype)
    else:
        val = rng.standard_normal(len(idx)).astype(dtype)
    return idx, val


class Recorder:
    def __init__(self):
        self.arrays = {}
        self.manifest = []

    def add(self, name, op, params, inputs, outputs):
        assert name not in [c["name"] for c in self.manifest], name
        for k, v in inputs.items():
            self.arrays["%s/in/%s" % (name, k)] = np.asarray(v)
        for k, v in outputs.items():
            self.arrays["%s/out/%s" % (name, k)] = np.asarray(v)
        self.manifest.append({"name": name, "op": op, "params": params,
                              "in": sorted(inputs), "out": sorted(outputs)})


def both_paths(fn, numpy_only=False):
    """Evaluate fn(ref_pkg) on the Cython path and the numpy path; insist the
    outputs are identical; return the Cython-path outputs.  ``numpy_only`` is
    for the one input class that crashes the reference's Cython kernel (an
    empty slice divides by zero at _merge.pyx:58 under cdivision)."""
    outs = []
    for cy in ((False, False) if numpy_only else (True, False)):
        ref = ref_loader.load_reference(use_cython=cy)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            outs.append(fn(ref))
    a, b = outs
    assert sorted(a) == sorted(b)
    for k in a:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y, equal_nan=True), (k, x, y)
    return a


def sp_out(prefix, s):
    return {prefix + "indices": s.indices, prefix + "data": s.data,
            prefix + "shape": np.asarray(s.shape, dtype=np.int64)}


def main():
    rng = np.random.default_rng(20261019)
    rec = Recorder()

    # ---- kernel seam (compat.py names): KATs + random ---------------------
    kat = [
        ("kat", [0, 1, 4, 6, 8, 9], [2, 4, 5, 6, 7]),         # test_compat.py:10-18
        ("kat2", [0, 2, 4, 8], [1, 4, 6, 8]),                  # test_compat.py:20-29
        ("empty_a", [], [1, 2, 3]),
        ("empty_b", [5, 9], []),
        ("empty_both", [], []),
        ("identical", [3, 7, 11, 2 ** 40], [3, 7, 11, 2 ** 40]),
        ("disjoint_lo_hi", [0, 1, 2], [10, 11, 12, 13]),
        ("interleaved", list(range(0, 40, 2)), list(range(1, 41, 2))),
    ]
    for i in range(4):
        na, nb = int(rng.integers(0, 400)), int(rng.integers(0, 400))
        hi = int(rng.integers(50, 900))
        kat.append(("rand%d" % i,
                    np.unique(rng.integers(0, hi, na)).tolist(),
                    np.unique(rng.integers(0, hi, nb)).tolist()))
    for name, a, b in kat:
        a = np.asarray(a, dtype=np.int64)
        b = np.asarray(b, dtype=np.int64)

        def run(ref, a=a, b=b):
            c, lut, am, bm = ref.compat.union1d_sorted(a.copy(), b.copy(), return_masks=True)
            ic, ia, ib = ref.compat.intersect1d_sorted(a.copy(), b.copy(), return_inds=True)
            return {"union": c, "lut": np.asarray(lut, dtype=np.uint8), "a_only": am, "b_only": bm,
                    "inter": ic, "a_inds": ia, "b_inds": ib}
        rec.add("seam_" + name, "seam_merge", {}, {"a": a, "b": b}, both_paths(run))

    cr_cases = [
        ("kat_outer", [[0, 2, 1], [1, 2, 1], [1, 4, 2]], (2, 3, 4), 4, False),   # test_compat.py:38-46
        ("kat_inner", [[0, 2, 1], [1, 2, 1], [1, 4, 2]], (2, 3, 4), 2, True),
        ("row", [[3, 4, 1], [0, 7, 1]], (5, 7), 7, False),
        ("col", [[0, 5, 1], [2, 3, 1]], (5, 7), 5, False),
        ("box3", [[1, 4, 1], [0, 6, 2], [2, 5, 1]], (4, 6, 5), 27, False),
        ("diag", [[0, 4, 1], [1, 5, 1]], (4, 6), 4, True),
    ]
    for name, ranges, shape, size, inner in cr_cases:
        ranges = np.asarray(ranges, dtype=np.int64)

        def run(ref, ranges=ranges, shape=shape, size=size, inner=inner):
            return {"flat": np.asarray(ref.compat.combine_ranges(ranges.copy(), shape, size, inner=inner))}
        rec.add("seam_cr_" + name, "seam_combine_ranges",
                {"shape": list(shape), "size": size, "inner": inner},
                {"ranges": ranges}, both_paths(run))

    lr = [(0, 5, 1), (5, 1, 1), (5, 0, -1), (0, 5, -2), (2, 17, 3), (17, 2, -4), (3, 3, 1), (0, 1, 5)]

    def run(ref):
        return {"len": np.asarray([ref.compat.len_range(*t) for t in lr], dtype=np.int64)}
    rec.add("seam_len_range", "seam_len_range", {}, {"args": np.asarray(lr, dtype=np.int64)}, both_paths(run))

    # ---- pairwise binops (flat.py:409-434) --------------------------------
    shapes = [((37,), 0.4), ((13, 11), 0.3), ((5, 6, 7), 0.25), ((4, 3, 5, 6), 0.2), ((3, 4), 1.0)]
    dtype_pairs = [("float64", "float64"), ("float32", "float32"), ("int64", "int64"),
                   ("float32", "float64"), ("int64", "float32"), ("int32", "int32")]
    ops = ["add", "sub", "mul", "minimum", "maximum", "lt", "le", "gt", "ge", "eq", "ne"]
    for si, (shape, dens) in enumerate(shapes):
        for da, db in dtype_pairs:
            ai, av = rand_sparse(rng, shape, dens, da)
            bi, bv = rand_sparse(rng, shape, dens, db)
            if len(av) > 2 and av.dtype.kind == "f":
                av[0] = -0.0
                av[1] = 0.0
            for op in ops:
                def run(ref, op=op, ai=ai, av=av, bi=bi, bv=bv, shape=shape):
                    a = ref.FlatSparray(ai.copy(), av.copy(), shape=shape, is_canonical=True)
                    b = ref.FlatSparray(bi.copy(), bv.copy(), shape=shape, is_canonical=True)
                    r = {"add": lambda: a + b, "sub": lambda: a - b, "mul": lambda: a * b,
                         "minimum": lambda: a.minimum(b), "maximum": lambda: a.maximum(b),
                         "lt": lambda: a < b, "le": lambda: a <= b, "gt": lambda: a > b,
                         "ge": lambda: a >= b, "eq": lambda: a == b, "ne": lambda: a != b}[op]()
                    return sp_out("", r)
                rec.add("binop_s%d_%s_%s_%s" % (si, da, db, op), "binop",
                        {"op": op, "shape": list(shape)},
                        {"a_indices": ai, "a_data": av, "b_indices": bi, "b_data": bv},
                        both_paths(run))

    # ---- transpose (flat.py:97-112); shapes where the reference is right ---
    tcases = [((7, 5), None), ((7, 5), (1, 0)), ((3, 4, 5), None), ((3, 4, 5), (1, 0, 2)),
              ((3, 4, 5), (2, 0, 1)), ((3, 4, 5), (0, 2, 1)), ((2, 3, 4, 5), None),
              ((2, 3, 4, 5), (1, 3, 0, 2)), ((2, 3, 4, 5), (0, 1, 3, 2)), ((6, 1, 3), (2, 1, 0)),
              ((16, 8, 4), (2, 1, 0)), ((9,), None)]
    for ti, (shape, axes) in enumerate(tcases):
        for dt in ("float64", "float32", "int64"):
            ai, av = rand_sparse(rng, shape, 0.4, dt)

            def run(ref, ai=ai, av=av, shape=shape, axes=axes):
                a = ref.FlatSparray(ai.copy(), av.copy(), shape=shape, is_canonical=True)
                r = a.transpose() if axes is None else a.transpose(*axes)
                return sp_out("", r)
            rec.add("transpose_%d_%s" % (ti, dt), "transpose",
                    {"shape": list(shape), "axes": None if axes is None else list(axes)},
                    {"indices": ai, "data": av}, both_paths(run))

    # ---- sum / mean over an axis (flat.py:607-626, base.py:84-100) --------
    scases = [((23,), [None, 0]), ((9, 7), [None, 0, 1]), ((4, 5, 6), [0, 1, 2]),
              ((3, 4, 5, 6), [0, 1, 2, 3]), ((1, 8), [0, 1]), ((6, 1, 5), [1])]
    for si, (shape, axes) in enumerate(scases):
        for dt in ("float64", "float32", "int64", "int32"):
            ai, av = rand_sparse(rng, shape, 0.5, dt)
            for ax in axes:
                def run(ref, ai=ai, av=av, shape=shape, ax=ax):
                    a = ref.FlatSparray(ai.copy(), av.copy(), shape=shape, is_canonical=True)
                    out = {}
                    for nm, r in (("sum_", a.sum(axis=ax)), ("mean_", a.mean(axis=ax))):
                        if hasattr(r, "indices"):
                            out.update(sp_out(nm, r))
                        else:
                            out[nm + "scalar"] = np.asarray(r)
                    return out
                rec.add("reduce_%d_%s_ax%s" % (si, dt, ax), "reduce",
                        {"shape": list(shape), "axis": ax},
                        {"indices": ai, "data": av}, both_paths(run))

    # ---- dot (flat.py:538-568) --------------------------------------------
    dcases = [((12, 9), "float64"), ((12, 9), "float32"), ((5, 4, 6), "float64"), ((1, 7), "float64")]
    for di, (shape, dt) in enumerate(dcases):
        ai, av = rand_sparse(rng, shape, 0.4, dt)
        x = rng.standard_normal(shape[-1]).astype(dt)
        x[rng.random(shape[-1]) < 0.3] = 0

        def run(ref, ai=ai, av=av, shape=shape, x=x):
            a = ref.FlatSparray(ai.copy(), av.copy(), shape=shape, is_canonical=True)
            dense = a.dot(x)                                    # flat.py:568 (densifies)
            sp = a.dot(ref.FlatSparray.from_ndarray(x))         # flat.py:545-561 (scipy route)
            out = {"dense": np.asarray(dense)}
            out.update(sp_out("sp_", sp))
            return out
        rec.add("dot_%d" % di, "dot_vec", {"shape": list(shape)},
                {"indices": ai, "data": av, "x": x}, both_paths(run))
    ai, av = rand_sparse(rng, (40,), 0.5, "float64")
    x = rng.standard_normal(40)

    def run(ref):
        a = ref.FlatSparray(ai.copy(), av.copy(), shape=(40,), is_canonical=True)
        return {"dense": np.asarray(a.dot(x)),
                "sp": np.asarray(a.dot(ref.FlatSparray.from_ndarray(x)))}
    rec.add("dot_1d", "dot_1d", {"shape": [40]}, {"indices": ai, "data": av, "x": x}, both_paths(run))

    # ---- __getitem__ (flat.py:254-325) -------------------------------------
    # (queries with repeated fancy indices are left out: the reference's two
    # kernel paths disagree on them -- intersect1d_sorted assumes unique inputs;
    # likewise inner-only fancy queries must ravel to an ascending sequence,
    # flat.py:278-280 never sorts them)
    gshape = (6, 5, 7)
    ai, av = rand_sparse(rng, gshape, 0.5, "float64")
    gcases = {
        "int_all": "(2, 3, 4)", "int_all_neg": "(-1, -2, -3)", "int_absent": "(0, 0, 0)",
        "row": "(3,)", "row_neg": "(-2,)", "partial": "(1, 2)", "col": "(slice(None), 2)",
        "ell_last": "(Ellipsis, 3)", "ell_mid": "(1, Ellipsis, 2)",
        "slice_box": "(slice(1, 5), slice(None), slice(2, 6))",
        "slice_step": "(slice(0, 6, 2), slice(1, 5, 3), slice(None, None, 2))",
        "slice_int_mix": "(slice(2, None), 1, slice(None, 4))",
        "empty_slice": "(slice(3, 3),)",
        "fancy_inner": "([0, 2, 5], [1, 1, 4], [6, 0, 3])",
        "fancy_slice": "([1, 3], slice(None), slice(2, 5))",
        "fancy_int": "([4, 0, 3], 2, slice(None))",
        "bool_1d": "(np.array([True, False, True, True, False, False]),)",
    }
    for name, expr in gcases.items():
        def run(ref, expr=expr):
            a = ref.FlatSparray(ai.copy(), av.copy(), shape=gshape, is_canonical=True)
            r = a[eval(expr)]
            if hasattr(r, "indices"):
                return sp_out("", r)
            return {"scalar": np.asarray(r)}
        rec.add("getitem_" + name, "getitem", {"shape": list(gshape), "index": expr},
                {"indices": ai, "data": av}, both_paths(run, numpy_only=(name == "empty_slice")))
    g1, v1 = rand_sparse(rng, (50,), 0.5, "int64")
    for name, expr in {"slice": "(slice(5, 40, 3),)", "fancy": "([0, 3, 17, 21, 49],)",
                       "int": "(7,)", "neg": "(-1,)"}.items():
        def run(ref, expr=expr):
            a = ref.FlatSparray(g1.copy(), v1.copy(), shape=(50,), is_canonical=True)
            r = a[eval(expr)]
            if hasattr(r, "indices"):
                return sp_out("", r)
            return {"scalar": np.asarray(r)}
        rec.add("getitem1d_" + name, "getitem", {"shape": [50], "index": expr},
                {"indices": g1, "data": v1}, both_paths(run))

    # ---- unary / scalar maps (flat.py:580-581,677-739) ---------------------
    ai, av = rand_sparse(rng, (8, 9), 0.5, "float64")
    av[:3] = [-0.0, 0.0, -2.5]
    ucases = {"neg": "-a", "abs": "abs(a)", "mul3": "a * 3", "mulf": "a * -3.5", "mul0": "a * 0",
              "div3": "a / 3", "pow2": "a ** 2", "pow3": "a ** 3", "f32": "a.astype(np.float32)",
              "i64": "a.astype(np.int64)", "sin": "a.sin()", "tan": "a.tan()", "sign": "a.sign()",
              "floor": "a.floor()", "ceil": "a.ceil()", "trunc": "a.trunc()", "rint": "a.rint()",
              "expm1": "a.expm1()", "arctan": "a.arctan()", "sinh": "a.sinh()", "tanh": "a.tanh()",
              "arcsinh": "a.arcsinh()", "deg2rad": "a.deg2rad()", "rad2deg": "a.rad2deg()",
              "copy": "a.copy()", "lt0": "a < -1", "gt1": "a > 1", "ne0": "a != 0",
              "min2": "a.minimum(2)", "max_m1": "a.maximum(-1)", "floordiv": "a // 2"}
    for name, expr in ucases.items():
        def run(ref, expr=expr):
            a = ref.FlatSparray(ai.copy(), av.copy(), shape=(8, 9), is_canonical=True)
            return sp_out("", eval(expr))
        rec.add("unary_" + name, "unary", {"shape": [8, 9], "expr": expr},
                {"indices": ai, "data": av}, both_paths(run))
    pos = np.abs(av)
    for name, expr in {"sqrt": "a.sqrt()", "log1p": "a.log1p()", "arcsin": "(a * 0.1).arcsin()",
                       "arctanh": "(a * 0.1).arctanh()"}.items():
        def run(ref, expr=expr):
            a = ref.FlatSparray(ai.copy(), pos.copy(), shape=(8, 9), is_canonical=True)
            return sp_out("", eval(expr))
        rec.add("unary_" + name, "unary", {"shape": [8, 9], "expr": expr},
                {"indices": ai, "data": pos}, both_paths(run))

    # ---- non-canonical ctor (flat.py:32-35) --------------------------------
    for dt in ("float64", "float32", "int64"):
        idx = rng.integers(0, 60, size=80).astype(np.int64)
        val = rand_sparse(rng, (200,), 1.0, dt)[1][:80]

        def run(ref, idx=idx, val=val):
            return sp_out("", ref.FlatSparray(idx.copy(), val.copy(), shape=(6, 10)))
        rec.add("ctor_%s" % dt, "ctor", {"shape": [6, 10]}, {"indices": idx, "data": val}, both_paths(run))

    rec.arrays["manifest"] = np.asarray(json.dumps(rec.manifest))
    np.savez_compressed(OUT, **rec.arrays)
    print("wrote %s: %d cases, %.1f KB" % (OUT, len(rec.manifest), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
