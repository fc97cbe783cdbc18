import cProfile, pstats, sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
import sparray_b200 as sp
ai, av, bi, bv = bench.gen_operands(0)
A = sp.FlatSparray(ai, av, shape=bench.SHAPE, is_canonical=True)
B = sp.FlatSparray(bi, bv, shape=bench.SHAPE, is_canonical=True)
def step():
    c = A * B
    return c.sum(axis=2)
for _ in range(20): step()
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter()-t0)/200*1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
