"""Host enqueue time vs device time of the bench step (c = a*b; c.sum(axis=2)) with lazily counted results."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import bench
import sparray_b200 as sp
ai, av, bi, bv = bench.gen_operands(0)
A = sp.FlatSparray(ai, av, shape=bench.SHAPE, is_canonical=True)
B = sp.FlatSparray(bi, bv, shape=bench.SHAPE, is_canonical=True)
def step():
    c = A * B
    return c.sum(axis=2)
for _ in range(300): step()
torch.cuda.synchronize()
for n in (50, 200, 1000):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): s = step()
    t1 = time.perf_counter(); e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("n=%d host enqueue %.1f us/step, wall %.1f us/step, device events %.1f us/step" % (n, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6, e0.elapsed_time(e1) / n * 1e3))
# device-only: same kernels through the C ABI with preallocated outputs
from sparray_b200 import _kernels as k
