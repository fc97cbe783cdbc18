"""Generate tests/golden/golden_f_v1.npz by RUNNING THE REFERENCE ITSELF: the SURVEY.md section 8(f) rows
(__setitem__ / setdiag / diagonal, converters, sparse broadcasting, dense (op) sparse, sparse . sparse dot,
the dense-returning branches).  Same rules as make_golden.py: every case runs on the reference's Cython kernel
path and on its numpy fallback, and both must agree bit for bit.

    python tests/golden/make_golden_f.py          (build container only: needs /root/reference)

A case is a small PROGRAM over the public API: ``setup`` statements executed with ``a`` (and ``b``) bound to
FlatSparray objects, ``x`` / ``d`` to dense arrays, then ``result`` is evaluated.  Results are sparse
(indices, data, shape) or dense arrays.  Programs that hit a scipy.sparse API removed from the installed scipy
(spmatrix.A) are phrased without it.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
from make_golden import Recorder, both_paths, rand_sparse  # noqa: E402

OUT = os.path.join(HERE, "golden_f_v1.npz")


def result_out(r):
    if hasattr(r, "indices") and hasattr(r, "data") and not isinstance(r, np.ndarray):
        return {"indices": np.asarray(r.indices), "data": np.asarray(r.data), "shape": np.asarray(r.shape, dtype=np.int64)}
    if hasattr(r, "tocoo") and not hasattr(r, "indices"):          # scipy matrix
        r = r.toarray()
    return {"dense": np.asarray(r)}


def main():
    rng = np.random.default_rng(20261020)
    rec = Recorder()

    def add(name, shape, dtype, setup, result, density=0.4, b_shape=None, extra=None):
        ai, av = rand_sparse(rng, shape, density, dtype)
        inputs = {"a_indices": ai, "a_data": av}
        params = {"shape": list(shape), "setup": setup, "result": result}
        if b_shape is not None:
            bi, bv = rand_sparse(rng, b_shape, density, dtype)
            inputs.update({"b_indices": bi, "b_data": bv})
            params["b_shape"] = list(b_shape)
        for k, v in (extra or {}).items():
            inputs[k] = np.asarray(v)

        def run(ref):
            env = {"np": np, "a": ref.FlatSparray(ai.copy(), av.copy(), shape=shape, is_canonical=True),
                   "FlatSparray": ref.FlatSparray}
            if b_shape is not None:
                env["b"] = ref.FlatSparray(inputs["b_indices"].copy(), inputs["b_data"].copy(), shape=b_shape, is_canonical=True)
            for k in (extra or {}):
                env[k] = np.array(inputs[k], copy=True)
            for st in setup:
                exec(st, env)
            return result_out(eval(result, env))
        rec.add(name, "prog", params, inputs, both_paths(run))

    # ---- f1: __setitem__ (flat.py:327-364, 395-407), setdiag (:139-160), diagonal (:114-137) ------------
    sh = (7, 9)
    for i, st in enumerate([
            "a[2, 3] = 5.0", "a[0, 0] = 0.0", "a[-1, -2] = -1.5", "a[6, 8] = 2.0",
            "a[1] = 7.5", "a[:, 2] = -1.0", "a[1:4, ::2] = 2.0", "a[2:3] = 0.0", "a[-2:, 3:] = 4.0",
            "a[3] = np.arange(9.0)", "a[:, 4] = np.arange(7.0) - 3", "a[1:3, 2:5] = np.arange(6.0)"]):
        add("setitem_%d" % i, sh, "float64", [st], "a")
    add("setitem_int64", sh, "int64", ["a[2, 3] = 5", "a[:, 1] = 4"], "a")
    add("setitem_3d", (3, 4, 5), "float64", ["a[1, 2] = 3.0", "a[2, :, 1] = -2.0", "a[0, 1, 2] = 9.0"], "a")
    add("setitem_1d", (30,), "float64", ["a[5:20:3] = 1.0", "a[29] = 2.0"], "a")
    for off in (0, 1, 3, -1, -4):
        add("setdiag_%d" % off, sh, "float64", ["a.setdiag(np.arange(1.0, 8.0)[:min(7 + min(%d, 0), 9 - max(%d, 0))], %d)" % (off, off, off)], "a")
        add("diagonal_%d" % off, sh, "float64", [], "a.diagonal(%d)" % off, density=0.6)
    add("setdiag_scalar", (6, 6), "float64", ["a.setdiag(2.5)"], "a")

    # ---- f2: converters (flat.py:48-82) ---------------------------------------------------------------------
    for i, (shape, dt) in enumerate([((8, 9), "float64"), ((4, 5, 6), "float32"), ((17,), "int64")]):
        add("toarray_%d" % i, shape, dt, [], "a.toarray()")
        add("from_ndarray_%d" % i, shape, dt, ["d = a.toarray()", "d.flat[::7] = 0"], "FlatSparray.from_ndarray(d)")
    add("tocoo", (8, 9), "float64", [], "a.tocoo().toarray()")
    add("from_spmatrix", (8, 9), "float64", ["import scipy.sparse as ss", "m = ss.csr_matrix(a.toarray())"], "FlatSparray.from_spmatrix(m)")

    # ---- f3: broadcasting between sparse operands (flat.py:451-472) -------------------------------------------
    add("bcast_add_col", (6, 7), "float64", [], "a + b", b_shape=(6, 1))
    add("bcast_add_row", (6, 7), "float64", [], "a + b", b_shape=(1, 7))
    add("bcast_mul_col", (6, 7), "float64", [], "a * b", b_shape=(6, 1))
    add("bcast_sub_row", (6, 7), "float64", [], "a - b", b_shape=(1, 7))
    add("bcast_3d", (3, 4, 5), "float64", [], "a + b", b_shape=(1, 4, 1))

    # ---- f4: dense (op) sparse (flat.py:436-449), sparse . sparse dot (:545-561) ---------------------------------
    dd = rng.standard_normal((6, 7))
    add("dense_add", (6, 7), "float64", [], "a + x", extra={"x": dd})
    add("dense_radd", (6, 7), "float64", [], "x + a", extra={"x": dd})
    add("dense_sub", (6, 7), "float64", [], "a - x", extra={"x": dd})
    add("dense_mul", (6, 7), "float64", [], "a * x", extra={"x": dd})
    add("dense_rmul", (6, 7), "float64", [], "x * a", extra={"x": dd})
    add("dense_mul_bcast", (6, 7), "float64", [], "a * x", extra={"x": dd[:, :1]})
    add("dense_min", (6, 7), "float64", [], "a.minimum(x)", extra={"x": dd})
    add("dense_max", (6, 7), "float64", [], "a.maximum(x)", extra={"x": dd})
    add("spgemm_0", (8, 6), "float64", [], "a.dot(b)", b_shape=(6, 9))
    add("spgemm_1", (5, 12), "float64", [], "a.dot(b)", b_shape=(12, 4), density=0.2)
    add("spgemm_vec", (8, 6), "float64", [], "a.dot(b)", b_shape=(6,))
    add("dot_dense_matrix", (8, 6), "float64", [], "a.dot(x)", extra={"x": rng.standard_normal((6, 3))})

    # ---- dense-returning branches (flat.py:478-481, 494-496, 520-527, 570-577) ---------------------------------
    add("scalar_add", (6, 7), "float64", [], "a + 1.5")
    add("scalar_radd", (6, 7), "float64", [], "2 + a")
    add("scalar_sub", (6, 7), "float64", [], "a - 1")
    add("scalar_lt", (6, 7), "float64", [], "a < 1")
    add("scalar_ge", (6, 7), "float64", [], "a >= 0")
    add("scalar_eq0", (6, 7), "float64", [], "a == 0")
    add("pow0", (6, 7), "float64", [], "a ** 0")

    # ---- complex values: the reference's own complex test data is ``dense * 1j`` (test_math.py:18-20) --------------
    cz = ["c = a * 1j"]
    add("cplx_times_i", (6, 7), "float64", [], "a * 1j")
    add("cplx_neg", (6, 7), "float64", cz, "-c")
    add("cplx_conj", (6, 7), "float64", cz, "c.conj()")
    add("cplx_scalar", (6, 7), "float64", cz, "c * 3")
    add("cplx_add_real", (6, 7), "float64", cz, "c + b", b_shape=(6, 7))
    add("cplx_add_cplx", (6, 7), "float64", cz + ["e = b * 1j"], "c + e", b_shape=(6, 7))
    add("cplx_sub_cplx", (6, 7), "float64", cz + ["e = b * 1j"], "c - e", b_shape=(6, 7))
    add("cplx_mul_real", (6, 7), "float64", cz, "c * b", b_shape=(6, 7))
    add("cplx_abs", (6, 7), "float64", cz, "abs(c)")
    add("cplx_getitem", (6, 7), "float64", cz, "c[2:5, 1::2]")
    add("cplx_toarray", (6, 7), "float64", cz, "c.toarray()")
    add("cplx_from_ndarray", (6, 7), "float64", ["d = a.toarray() * (2 - 1j)"], "FlatSparray.from_ndarray(d)")
    add("cplx_sum", (6, 7), "float64", cz, "np.asarray(c.sum())")

    rec.arrays["manifest"] = np.asarray(json.dumps(rec.manifest))
    np.savez_compressed(OUT, **rec.arrays)
    print("wrote %s: %d cases, %.1f KB" % (OUT, len(rec.manifest), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
