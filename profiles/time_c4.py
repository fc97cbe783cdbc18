"""Event-timed API-level medians of the config-4 ops (1e6 x 1e6 float64, 100 M stored elements): a.dot(x), a.sum(axis=1).
Usage: python profiles/time_c4.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sparray_b200 as sp
from run_kernels import synth

def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts) * 1e3

wi, wv = synth(100_000_000, 10 ** 12, torch.float64, 4)
W = sp.FlatSparray(wi, wv, shape=(10 ** 6, 10 ** 6), is_canonical=True)
x = torch.randn(10 ** 6, device="cuda", dtype=torch.float64)
xn = x.cpu().numpy()
print("c4 sum(axis=1) f64 100M: %.0f us" % t(lambda: W.sum(axis=1)))
from sparray_b200 import _kernels as k
print("c4 spmv kernel f64 100M: %.0f us" % t(lambda: k.spmv(wi, wv, 10 ** 6, 10 ** 6, x, torch.float64)))
W32 = sp.FlatSparray(wi, wv.float(), shape=(10 ** 6, 10 ** 6), is_canonical=True)
print("c4 sum(axis=1) f32 100M: %.0f us" % t(lambda: W32.sum(axis=1)))
