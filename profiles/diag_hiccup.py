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
def run(n, tag):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    hs = []
    evs[0].record()
    for i in range(n):
        t0 = time.perf_counter()
        s = step()
        hs.append((time.perf_counter() - t0) * 1e6)
        evs[i + 1].record()
    torch.cuda.synchronize()
    d = [evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(n)]
    print(tag, "total %.2f ms; device max %.0f us at %d; host max %.0f us at %d; host median %.0f" % (sum(d) / 1e3, max(d), d.index(max(d)), max(hs), hs.index(max(hs)), sorted(hs)[n // 2]))
for _ in range(20): step()
torch.cuda.synchronize()
run(20, "no sampler A")
run(20, "no sampler B")
smp = bench.ClockSampler(0)
smp.start(); smp.wait_ready()
run(20, "sampler on A")
run(20, "sampler on B")
time.sleep(0.1)
run(20, "sampler on C")
run(200, "sampler on D")
print(smp.stop())
run(20, "after stop")
