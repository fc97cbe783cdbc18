"""Event-timed API-level medians of the non-merge ops on config-shaped operands.
Usage: python profiles/time_ops.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sparray_b200 as sp
from sparray_b200 import _kernels as k
from run_kernels import synth

def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts) * 1e3

shape = (256, 256, 256, 64)
ai, av = synth(10_700_000, 256 * 256 * 256 * 64, torch.float32, 1)
A = sp.FlatSparray(ai, av, shape=shape, is_canonical=True)
print("sum_axis3 (last) f32 10.7M: %.0f us" % t(lambda: A.sum(axis=3)))
print("sum_axis2 f32 10.7M: %.0f us" % t(lambda: A.sum(axis=2)))
for ax in (3, 2, 1):
    g = k.slab_geometry(shape, ax)
    if sp._lib.lib.spb_axis_sum_slab_ok(ai.numel(), *g):
        print("  slab kernels, axis %d: %.0f us" % (ax, t(lambda: k.axis_sum_slab(ai, av, g, ai.numel()))))
print("  reduce_by_key(div 64) = last-axis sum without the dense slab: %.0f us" % t(lambda: k.reduce_by_key(ai, av, 64)))
print("  sum_axis1 f32 10.7M: %.0f us ; sum_axis0: %.0f us" % (t(lambda: A.sum(axis=1)), t(lambda: A.sum(axis=0))))
print("transpose 3210 f32 10.7M: %.0f us" % t(lambda: A.transpose(3, 2, 1, 0)))
wi, wv = synth(100_000_000, 10 ** 12, torch.float64, 4)
W = sp.FlatSparray(wi, wv, shape=(10 ** 6, 10 ** 6), is_canonical=True)
print("sum_axis1 f64 100M: %.0f us" % t(lambda: W.sum(axis=1), reps=5))
x = torch.randn(10 ** 6, device="cuda", dtype=torch.float64)
print("spmv f64 100M: %.0f us" % t(lambda: k.spmv(W.indices, W.data, 10 ** 6, 10 ** 6, x, torch.float64), reps=5))
xs = torch.randn(64, device="cuda", dtype=torch.float32)
print("spmv f32 10.7M, 64 columns (0.64 elements per row): %.0f us" % t(lambda: k.spmv(ai, av, 64, 256 ** 3, xs, torch.float32)))
