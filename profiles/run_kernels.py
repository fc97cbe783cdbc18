"""Small driver used for ncu captures: launches each hot kernel a few times on
config-shaped synthetic operands through the C ABI (no timing here; numbers taken
under a profiler are never bench values).  Usage: python profiles/run_kernels.py [op ...]
ops: intersect union sum_last sum_mid transpose spmv sum_c4"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparray_b200 as sp  # noqa: E402
from sparray_b200 import _kernels as k  # noqa: E402


def synth(n, size, dtype, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    gap = max(size // n, 1)
    steps = torch.randint(1, 2 * gap, (n,), device="cuda", generator=g, dtype=torch.int64)
    idx = torch.cumsum(steps, 0) - 1
    idx = idx[idx < size].contiguous()
    return idx, torch.randn(idx.numel(), device="cuda", generator=g, dtype=dtype)


def main():
    ops = sys.argv[1:] or ["intersect", "union", "sum_last", "sum_mid", "transpose", "spmv"]
    shape = (256, 256, 256, 64)
    size = 256 * 256 * 256 * 64
    ai, av = synth(10_700_000, size, torch.float32, 1)
    bi, bv = synth(10_700_000, size, torch.float32, 2)
    A = sp.FlatSparray(ai, av, shape=shape, is_canonical=True)
    B = sp.FlatSparray(bi, bv, shape=shape, is_canonical=True)
    reps = 3
    for _ in range(reps):
        if "intersect" in ops:
            A * B
        if "union" in ops:
            A + B
        if "sum_last" in ops:
            A.sum(axis=3)
        if "sum_mid" in ops:
            A.sum(axis=2)
        if "transpose" in ops:
            A.transpose(3, 2, 1, 0)
    if "spmv" in ops:
        wi, wv = synth(100_000_000, 10 ** 12, torch.float64, 4)
        x = torch.randn(10 ** 6, device="cuda", dtype=torch.float64)
        for _ in range(reps):
            k.spmv(wi, wv, 10 ** 6, 10 ** 6, x, torch.float64)
    if "sum_c4" in ops:
        wi, wv = synth(100_000_000, 10 ** 12, torch.float64, 4)
        W = sp.FlatSparray(wi, wv, shape=(10 ** 6, 10 ** 6), is_canonical=True)
        for _ in range(reps):
            W.sum(axis=1)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
