"""Randomised stress run against numpy on dense arrays (stands in for the sanitizer runs this pool refuses):
random shapes / fills / dtypes, every op twice in a row (a dirty workspace, stale look-back words or counters
would show up on the second call).  Usage: python profiles/stress_parity.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sparray_b200 as sp

def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    t0 = time.time(); it = 0; checks = 0
    while time.time() - t0 < budget:
        it += 1
        ndim = int(rng.integers(1, 5))
        shape = tuple(int(x) for x in rng.integers(1, [400, 90, 40, 12][:ndim] if ndim > 1 else [200000]))
        size = int(np.prod(shape))
        dens = float(rng.choice([0.001, 0.01, 0.1, 0.5, 0.95]))
        dtype = rng.choice([np.float64, np.float32, np.int64, np.int32])
        def dense():
            d = np.zeros(shape, dtype=dtype)
            m = rng.random(shape) < dens
            k = int(m.sum())
            d[m] = (rng.standard_normal(k) * 10).astype(dtype) if np.dtype(dtype).kind == "f" else rng.integers(-50, 50, k).astype(dtype)
            return d
        da, db = dense(), dense()
        A, B = sp.FlatSparray.from_ndarray(da), sp.FlatSparray.from_ndarray(db)
        for rep in range(2):
            for name, got, want in (("add", A + B, da + db), ("mul", A * B, da * db), ("sub", A - B, da - db),
                                    ("min", A.minimum(B), np.minimum(da, db)), ("lt", A < B, da < db)):
                g = got.toarray()
                assert g.dtype == want.dtype and np.array_equal(g, want), (name, shape, dens, dtype, rep)
                gi = got.indices.cpu().numpy()
                assert np.all(np.diff(gi) > 0), ("unsorted / duplicate indices", name, shape)
                checks += 1
            for ax in range(ndim):
                g = A.sum(axis=ax).toarray() if ndim > 1 else np.asarray(A.sum(axis=ax))
                w = da.astype(np.float64).sum(axis=ax)
                tol = 1e-4 if dtype == np.float32 else 1e-10
                assert np.allclose(g, w, rtol=tol, atol=tol * max(1.0, np.abs(w).max())), ("sum", ax, shape, dens, dtype, rep)
                checks += 1
            if ndim > 1:
                perm = tuple(int(x) for x in rng.permutation(ndim))
                g = A.transpose(*perm)
                assert np.array_equal(g.toarray(), da.transpose(perm)), ("transpose", perm, shape, dens, dtype, rep)
                gi = g.indices.cpu().numpy()
                assert np.all(np.diff(gi) > 0), ("transpose unsorted", perm, shape)
                checks += 1
            if ndim == 2:
                x = rng.standard_normal(shape[1])
                g = np.asarray(A.dot(x)) if not hasattr(A.dot(x), "cpu") else A.dot(x).cpu().numpy()
                w = da.astype(np.float64).dot(x)
                assert np.allclose(g, w, rtol=1e-4 if dtype == np.float32 else 1e-10, atol=1e-6 * max(1.0, np.abs(w).max())), ("dot", shape, dens, dtype)
                checks += 1
    torch.cuda.synchronize()
    print("stress ok: %d random arrays, %d checks in %.0f s (seed %d)" % (it, checks, time.time() - t0, seed))

if __name__ == "__main__":
    main()
