"""Kernel-only timing of the fused setop calls (CUDA events around the C-ABI call, config-2
operands).  Usage: python profiles/time_setop.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparray_b200 as sp
from sparray_b200 import _dtypes as dt
from sparray_b200._lib import check, lib, ptr, stream, workspace
sys.path.insert(0, os.path.join(os.path.dirname(__file__)))
from run_kernels import synth

def main():
  for tdt, sdt, tag in ((torch.float32, dt.SPB_F32, "f32"), (torch.float64, dt.SPB_F64, "f64")):
    size = 256 * 256 * 256 * 64
    ai, av = synth(10_700_000, size, tdt, 1)
    bi, bv = synth(10_700_000, size, tdt, 2)
    na, nb = ai.numel(), bi.numel()
    out_idx = torch.empty(na + nb, dtype=torch.int64, device="cuda")
    out_val = torch.empty(na + nb, dtype=tdt, device="cuda")
    cnt = torch.empty(1, dtype=torch.int64, device="cuda")
    for name, fn, op in (("intersect " + tag, lib.spb_intersect_binop, "mul"), ("union " + tag, lib.spb_union_binop, "add")):
        def launch():
            check(fn(dt.BINOPS[op], sdt, ptr(ai), ptr(av), na, ptr(bi), ptr(bv), nb, ptr(out_idx), ptr(out_val),
                     ptr(cnt), workspace(), stream()))
        for _ in range(5):
            launch()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); launch(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print("%s: median %.1f us  min %.1f us  (nout=%d)" % (name,
              np.median(ts) * 1e3, np.min(ts) * 1e3, int(cnt.item())))

if __name__ == "__main__":
    main()
