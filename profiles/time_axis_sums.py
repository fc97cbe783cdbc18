import sys; sys.path.insert(0, "profiles"); sys.path.insert(0, ".")
import torch, numpy as np, sparray_b200 as sp
from run_kernels import synth
ai, av = synth(10_700_000, 256*256*256*64, torch.float32, 1)
A = sp.FlatSparray(ai, av, shape=(256,256,256,64), is_canonical=True)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts) * 1e3
print("axis0 %.0f us  axis1 %.0f us  axis2 %.0f us" % (t(lambda: A.sum(axis=0)), t(lambda: A.sum(axis=1)), t(lambda: A.sum(axis=2))))
