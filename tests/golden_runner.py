"""Evaluate the committed golden cases (tests/golden/golden_v1.npz, produced by
running the reference: tests/golden/make_golden.py) on any backend that offers
the FlatSparray surface -- the CPU oracle or the CUDA product."""
import json
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = os.path.join(_DIR, "golden_v1.npz")
GOLDEN_F = os.path.join(_DIR, "golden_f_v1.npz")      # make_golden_f.py
GOLDEN_G = os.path.join(_DIR, "golden_g_v1.npz")      # long rows (last-axis sums, dot): make_golden_g.py
_CACHE = {}


def load():
    if "z" not in _CACHE:
        arrays, manifest = {}, []
        for path in (GOLDEN, GOLDEN_F, GOLDEN_G):
            z = np.load(path)
            arrays.update({k: z[k] for k in z.files if k != "manifest"})
            manifest += json.loads(str(z["manifest"]))
        _CACHE["z"] = arrays
        _CACHE["manifest"] = manifest
    return _CACHE["manifest"], _CACHE["z"]


def cases(op=None):
    manifest, _ = load()
    return [c for c in manifest if op is None or c["op"] == op]


def arrays(case, which):
    _, z = load()
    return {k: z["%s/%s/%s" % (case["name"], which, k)] for k in case[which]}


def to_np(x):
    """Host copy of a numpy array, torch tensor, memoryview or scalar."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def sp_out(prefix, s):
    return {prefix + "indices": to_np(s.indices), prefix + "data": to_np(s.data),
            prefix + "shape": np.asarray(tuple(s.shape), dtype=np.int64)}


def run_case(case, cls, seam=None):
    """cls: FlatSparray-like class.  seam: object with union1d_sorted,
    intersect1d_sorted, combine_ranges, len_range (for the seam_* cases)."""
    op, prm, inp = case["op"], case["params"], arrays(case, "in")
    if op == "seam_merge":
        c, lut, am, bm = seam.union1d_sorted(seam.asinput(inp["a"]), seam.asinput(inp["b"]), return_masks=True)
        ic, ia, ib = seam.intersect1d_sorted(seam.asinput(inp["a"]), seam.asinput(inp["b"]), return_inds=True)
        plain_u = seam.union1d_sorted(seam.asinput(inp["a"]), seam.asinput(inp["b"]))
        plain_i = seam.intersect1d_sorted(seam.asinput(inp["a"]), seam.asinput(inp["b"]))
        assert np.array_equal(to_np(plain_u), to_np(c)) and np.array_equal(to_np(plain_i), to_np(ic))
        return {"union": to_np(c), "lut": to_np(lut).astype(np.uint8), "a_only": to_np(am).astype(bool),
                "b_only": to_np(bm).astype(bool), "inter": to_np(ic), "a_inds": to_np(ia), "b_inds": to_np(ib)}
    if op == "seam_combine_ranges":
        return {"flat": to_np(seam.combine_ranges(inp["ranges"], tuple(prm["shape"]), prm["size"],
                                                  inner=prm["inner"]))}
    if op == "seam_len_range":
        return {"len": np.asarray([seam.len_range(*map(int, t)) for t in inp["args"]], dtype=np.int64)}
    shape = tuple(prm["shape"])
    if op == "prog":
        # a small program over the public API (tests/golden/make_golden_f.py)
        env = {"np": np, "FlatSparray": cls, "a": cls(inp["a_indices"], inp["a_data"], shape=shape, is_canonical=True)}
        if "b_indices" in inp:
            env["b"] = cls(inp["b_indices"], inp["b_data"], shape=tuple(prm["b_shape"]), is_canonical=True)
        for k, v in inp.items():
            if k not in ("a_indices", "a_data", "b_indices", "b_data"):
                env[k] = np.array(v, copy=True)
        for st in prm["setup"]:
            exec(st, env)
        r = eval(prm["result"], env)
        if hasattr(r, "indices") and hasattr(r, "shape") and not isinstance(r, np.ndarray) and not hasattr(r, "detach"):
            return sp_out("", r)
        if hasattr(r, "tocoo") and not hasattr(r, "indices"):
            r = r.toarray()
        return {"dense": to_np(r)}
    if op == "binop":
        a = cls(inp["a_indices"], inp["a_data"], shape=shape, is_canonical=True)
        b = cls(inp["b_indices"], inp["b_data"], shape=shape, is_canonical=True)
        r = {"add": lambda: a + b, "sub": lambda: a - b, "mul": lambda: a * b,
             "minimum": lambda: a.minimum(b), "maximum": lambda: a.maximum(b),
             "lt": lambda: a < b, "le": lambda: a <= b, "gt": lambda: a > b,
             "ge": lambda: a >= b, "eq": lambda: a == b, "ne": lambda: a != b}[prm["op"]]()
        return sp_out("", r)
    if op == "ctor":
        return sp_out("", cls(inp["indices"], inp["data"], shape=shape))
    a = cls(inp["indices"], inp["data"], shape=shape, is_canonical=True)
    if op == "transpose":
        r = a.transpose() if prm["axes"] is None else a.transpose(*prm["axes"])
        return sp_out("", r)
    if op == "reduce":
        out = {}
        for nm, r in (("sum_", a.sum(axis=prm["axis"])), ("mean_", a.mean(axis=prm["axis"]))):
            if hasattr(r, "indices"):
                out.update(sp_out(nm, r))
            else:
                out[nm + "scalar"] = to_np(r)
        return out
    if op == "dot_vec":
        return {"dense": to_np(a.dot(inp["x"]))}
    if op == "dot_1d":
        return {"dense": to_np(a.dot(inp["x"]))}
    if op == "getitem":
        r = a[eval(prm["index"], {"np": np, "Ellipsis": Ellipsis})]
        return sp_out("", r) if hasattr(r, "indices") else {"scalar": to_np(r)}
    if op == "unary":
        return sp_out("", eval(prm["expr"], {"np": np, "a": a, "abs": abs}))
    raise KeyError(op)


def check_case(case, got, rtol_f64=1e-12, rtol_f32=1e-5):
    """Bit-exact for indices/shape/bool/int and single elementwise float ops;
    the north_star tolerance for float reductions and dot."""
    want = arrays(case, "out")
    if case["op"] == "getitem" and "indices" in want:
        # flat.py:391-392 hands back indices in query order when the fancy query
        # was not ascending (a FlatSparray that is flagged canonical but is not
        # sorted); the product and the oracle return the sorted form.
        order = np.argsort(want["indices"], kind="stable")
        want = dict(want, indices=want["indices"][order], data=want["data"][order])
    exact = case["op"] in ("seam_merge", "seam_combine_ranges", "seam_len_range", "binop",
                           "transpose", "getitem", "ctor")
    if case["op"] == "prog":
        # everything but the contractions is data movement or one elementwise op: bit-exact
        exact = ".dot(" not in case["params"]["result"]
    for k, w in want.items():
        if k not in got:
            if case["op"] in ("dot_vec", "dot_1d") and k != "dense":
                continue            # sparse-route outputs are a second reference, not a product output
            raise AssertionError("%s: missing output %s" % (case["name"], k))
        g = got[k]
        msg = "%s/%s" % (case["name"], k)
        if k.endswith("shape"):
            assert tuple(g.ravel()) == tuple(w.ravel()), msg
            continue
        assert g.shape == w.shape, "%s shape %s != %s" % (msg, g.shape, w.shape)
        assert g.dtype == w.dtype, "%s dtype %s != %s" % (msg, g.dtype, w.dtype)
        if w.dtype.kind in "biu" or exact:
            same = np.array_equal(g, w, equal_nan=True) if w.dtype.kind in "fc" else np.array_equal(g, w)
            if same and w.dtype.kind == "f":
                same = np.array_equal(np.signbit(g), np.signbit(w))
            assert same, "%s\n got %r\nwant %r" % (msg, g, w)
        else:
            if case["op"] == "unary":
                rtol = 4 * np.finfo(w.dtype).eps     # libm vs CUDA math: a few ulp
            else:
                src = case["in"]
                f32 = any(arrays(case, "in")[s].dtype == np.float32 for s in src)
                rtol = rtol_f32 if f32 else rtol_f64
            scale = np.max(np.abs(w)) if w.size else 1.0
            np.testing.assert_allclose(g, w, rtol=rtol, atol=rtol * max(scale, 1e-300), err_msg=msg)
