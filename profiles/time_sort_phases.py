"""Phase timing of the radix pass kernel (needs a library built with -DSPB_SORT_TIMING, selected with
SPB_LIB_PATH): per-CTA global-timer stamps at the phase boundaries, averaged over the tiles of the LAST
pass of one sort.  Usage: SPB_LIB_PATH=variants/timing.so python profiles/time_sort_phases.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sparray_b200 as sp  # noqa: E402
from sparray_b200._lib import lib  # noqa: E402
from run_kernels import synth  # noqa: E402

NAMES = ["start->dep_sync", "dep_sync->loaded", "loaded->ranked", "barrier1", "digit sums+scan+(2)", "wait barrier3",
         "exchange (thread0)", "look-back (thread 0's digit)", "wait barrier4", "mbar wait (values)", "scatter"]


def main():
    shape = (256, 256, 256, 64)
    ai, av = synth(10_700_000, 256 * 256 * 256 * 64, torch.float32, 1)
    A = sp.FlatSparray(ai, av, shape=shape, is_canonical=True)
    for _ in range(3):
        A.transpose(3, 2, 1, 0)
    ntiles = 20000
    buf = torch.zeros(ntiles * 16, dtype=torch.int64, device="cuda")
    lib.spb_debug_sort_timing.restype = ctypes.c_int
    lib.spb_debug_sort_timing.argtypes = [ctypes.c_void_p]
    assert lib.spb_debug_sort_timing(ctypes.c_void_p(buf.data_ptr())) == 0
    A.transpose(3, 2, 1, 0)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(ntiles, 16)
    used = t[:, 0] > 0
    t = t[used][:, :12].astype(np.float64)
    print("tiles stamped (last pass): %d" % len(t))
    t0 = t[:, 0].min()
    print("kernel span (first stamp -> last stamp): %.1f us" % ((t[:, 11].max() - t0) / 1e3))
    d = np.diff(t, axis=1) / 1e3
    for i, nm in enumerate(NAMES):
        print("  %-32s mean %7.2f us   p50 %7.2f   p90 %7.2f" % (nm, d[:, i].mean(), np.median(d[:, i]), np.percentile(d[:, i], 90)))
    life = (t[:, 11] - t[:, 1]) / 1e3
    print("  tile lifetime after dep sync: mean %.2f us, p50 %.2f, p90 %.2f" % (life.mean(), np.median(life), np.percentile(life, 90)))
    # start order vs tile index
    order = np.argsort(t[:, 1], kind="stable")
    print("  tiles whose start order differs from index order by > 100: %d" % int((np.abs(order - np.arange(len(order))) > 100).sum()))
    st = (t[:, 1] - t0) / 1e3
    print("  start times (us) of tiles 0, 100, 500, 1000, 2000, last:", [round(float(st[i]), 1) for i in (0, 100, 500, 1000, min(2000, len(st) - 1), len(st) - 1)])


if __name__ == "__main__":
    main()
