"""Phase timing of the slab axis-sum kernel (needs a library built with -DSPB_SLAB_TIMING, selected with
SPB_LIB_PATH): per-CTA global-timer stamps at the phase boundaries.
Usage: SPB_LIB_PATH=variants/slabtiming.so python profiles/time_slab_phases.py [axis]"""
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

NAMES = ["start->dep_sync", "dep_sync->bounds+zeroed", "accumulate (thread 0)", "wait barrier A", "count + barrier B",
         "look-back (warp 0)", "list + barrier C", "emit"]


def main():
    axis = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    shape = (256, 256, 256, 64)
    ai, av = synth(10_700_000, 256 * 256 * 256 * 64, torch.float32, 1)
    A = sp.FlatSparray(ai, av, shape=shape, is_canonical=True)
    for _ in range(3):
        A.sum(axis=axis)
    ntiles = 20000
    buf = torch.zeros(ntiles * 16, dtype=torch.int64, device="cuda")
    lib.spb_debug_slab_timing.restype = ctypes.c_int
    lib.spb_debug_slab_timing.argtypes = [ctypes.c_void_p]
    assert lib.spb_debug_slab_timing(ctypes.c_void_p(buf.data_ptr())) == 0
    A.sum(axis=axis)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(ntiles, 16)
    used = t[:, 0] > 0
    t = t[used][:, :9].astype(np.float64)
    print("axis %d: tiles stamped: %d" % (axis, len(t)))
    t0 = t[:, 0].min()
    print("kernel span (first stamp -> last stamp): %.1f us" % ((t[:, 8].max() - t0) / 1e3))
    d = np.diff(t, axis=1) / 1e3
    for i, nm in enumerate(NAMES):
        print("  %-28s mean %7.2f us   p50 %7.2f   p90 %7.2f   max %7.2f" % (nm, d[:, i].mean(), np.median(d[:, i]), np.percentile(d[:, i], 90), d[:, i].max()))
    life = (t[:, 8] - t[:, 1]) / 1e3
    print("  tile lifetime after dep sync: mean %.2f us, p50 %.2f, p90 %.2f" % (life.mean(), np.median(life), np.percentile(life, 90)))
    st = (t[:, 1] - t0) / 1e3
    en = (t[:, 8] - t0) / 1e3
    idx = [0, 100, 500, 1000, 2000, 3000, 4000, len(st) - 1]
    print("  start times (us):", [round(float(st[min(i, len(st) - 1)]), 1) for i in idx])
    print("  end times   (us):", [round(float(en[min(i, len(st) - 1)]), 1) for i in idx])


if __name__ == "__main__":
    main()
