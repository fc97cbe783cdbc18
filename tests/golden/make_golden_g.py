"""Generate tests/golden/golden_g_v1.npz by RUNNING THE REFERENCE ITSELF: long rows -- sums over the last axis and
dot with a dense vector where an output cell collects tens to hundreds of stored elements, the shapes the
lane-owns-four kernels serve (csrc/axis_sum.cu axis_accumulate_lastaxis_kernel, csrc/reduce.cu spmv_runs_kernel).
Same rules as make_golden.py: every case runs on the reference's Cython kernel path and on its numpy fallback,
and both must agree bit for bit.  Case formats are make_golden.py's "reduce" and "dot_vec".

    python tests/golden/make_golden_g.py          (build container only: needs /root/reference)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import Recorder, both_paths, rand_sparse, sp_out  # noqa: E402

OUT = os.path.join(HERE, "golden_g_v1.npz")


def main():
    rng = np.random.default_rng(20261021)
    rec = Recorder()

    # ---- sum / mean over the last axis, long runs (flat.py:607-626, base.py:84-100) ----------------------------
    scases = [((30, 400), 0.3), ((6, 5, 300), 0.25), ((1, 2500), 0.4), ((300, 40), 0.5)]
    for si, (shape, dens) in enumerate(scases):
        for dt in ("float64", "float32", "int64", "int32"):
            ai, av = rand_sparse(rng, shape, dens, dt)
            # a few rows with a single stored element between the long ones (a lane of four then spans three rows)
            keep = np.ones(len(ai), dtype=bool)
            rows = ai // shape[-1]
            for r in np.unique(rows)[1::5]:
                sel = np.nonzero(rows == r)[0]
                keep[sel[1:]] = False
            ai, av = ai[keep], av[keep]
            ax = len(shape) - 1

            def run(ref, ai=ai, av=av, shape=shape, ax=ax):
                a = ref.FlatSparray(ai.copy(), av.copy(), shape=shape, is_canonical=True)
                out = {}
                for nm, r in (("sum_", a.sum(axis=ax)), ("mean_", a.mean(axis=ax))):
                    out.update(sp_out(nm, r))
                return out
            rec.add("reduce_long_%d_%s_ax%s" % (si, dt, ax), "reduce", {"shape": list(shape), "axis": ax},
                    {"indices": ai, "data": av}, both_paths(run))

    # ---- dot with a dense vector, long rows (flat.py:538-568) --------------------------------------------------
    dcases = [((30, 400), "float64", 0.3), ((30, 400), "float32", 0.3), ((6, 5, 300), "float64", 0.25),
              ((1, 2500), "float64", 0.4)]
    for di, (shape, dt, dens) in enumerate(dcases):
        ai, av = rand_sparse(rng, shape, dens, dt)
        x = rng.standard_normal(shape[-1]).astype(dt)
        x[rng.random(shape[-1]) < 0.3] = 0

        def run(ref, ai=ai, av=av, shape=shape, x=x):
            a = ref.FlatSparray(ai.copy(), av.copy(), shape=shape, is_canonical=True)
            dense = a.dot(x)                                    # flat.py:568 (densifies)
            sp = a.dot(ref.FlatSparray.from_ndarray(x))         # flat.py:545-561 (scipy route)
            out = {"dense": np.asarray(dense)}
            out.update(sp_out("sp_", sp))
            return out
        rec.add("dot_long_%d" % di, "dot_vec", {"shape": list(shape)},
                {"indices": ai, "data": av, "x": x}, both_paths(run))

    rec.arrays["manifest"] = np.asarray(json.dumps(rec.manifest))
    np.savez_compressed(OUT, **rec.arrays)
    print("wrote %s: %d cases, %.1f KB" % (OUT, len(rec.manifest), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
