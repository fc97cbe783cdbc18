"""Event-timed transposes / re-sorts at the BASELINE shapes (API level), with the traffic the passes
actually move, and the intersection call beside them.  Usage: python profiles/time_sort.py [big]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sparray_b200 as sp  # noqa: E402
from run_kernels import synth  # noqa: E402

PEAK = 6531.9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def t(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def report(name, n, v, ms):
    alg = 2 * n * (8 + v)
    print("%-44s n=%.3g  %.3f ms  %.1f Gnnz/s  alg %.0f GB/s = %.1f%% of HBM peak" % (
        name, n, ms, n / ms / 1e6, alg / ms / 1e6, 100 * alg / ms / 1e6 / PEAK), flush=True)


def main():
    shape = (256, 256, 256, 64)
    ai, av = synth(10_700_000, 256 * 256 * 256 * 64, torch.float32, 1)
    A = sp.FlatSparray(ai, av, shape=shape, is_canonical=True)
    report("c2 transpose(3,2,1,0) f32", A.nnz, 4, t(lambda: A.transpose(3, 2, 1, 0)))
    report("c2 swapaxes(0,2) f32", A.nnz, 4, t(lambda: A.swapaxes(0, 2)))
    report("c2 swapaxes(2,3) f32 (1 pass)", A.nnz, 4, t(lambda: A.swapaxes(2, 3)))
    zi, zv = synth(1_717_987, 2 ** 34, torch.float32, 3)
    Z = sp.FlatSparray(zi, zv, shape=(4096, 4096, 1024), is_canonical=True)
    report("c3 literal swapaxes(0,2) f32", Z.nnz, 4, t(lambda: Z.swapaxes(0, 2)))
    wi, wv = synth(100_000_000, 10 ** 12, torch.float64, 4)
    W = sp.FlatSparray(wi, wv, shape=(10 ** 6, 10 ** 6), is_canonical=True)
    report("c4 .T f64 100M", W.nnz, 8, t(lambda: W.T, reps=5, warm=2))
    del W, wi, wv
    torch.cuda.empty_cache()
    if "big" in sys.argv[1:]:
        i5, v5 = synth(1_717_986_918, 2 ** 34, torch.float32, 5)
        Zb = sp.FlatSparray(i5, v5, shape=(4096, 4096, 1024), is_canonical=True)
        del i5, v5
        report("c3 stated size swapaxes(0,2) f32 1.7G", Zb.nnz, 4, t(lambda: Zb.swapaxes(0, 2), reps=3, warm=1))


if __name__ == "__main__":
    main()
