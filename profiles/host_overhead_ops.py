"""Host-side cost of one public-API call (enqueue rate with the GPU never the limiter: tiny operands), and where
it goes (cProfile).  Usage: python profiles/host_overhead_ops.py"""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import sparray_b200 as sp
rng = np.random.default_rng(0)
shape = (1000, 1000)
def mk(seed):
    r = np.random.default_rng(seed)
    idx = np.unique(r.integers(0, 10**6, size=2000, dtype=np.int64))
    return sp.FlatSparray(idx, r.standard_normal(len(idx)), shape=shape, is_canonical=True)
A, B = mk(1), mk(2)
ops = {"a+b": lambda: A + B, "a*b": lambda: A * B, "a.sum(axis=1)": lambda: A.sum(axis=1), "a.T": lambda: A.T,
       "a*2.5": lambda: A * 2.5, "a[3]": lambda: A[3]}
for name, fn in ops.items():
    for _ in range(200): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("%-16s host %.1f us/call" % (name, (t1 - t0) / 2000 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(3000): A + B
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:4000])
