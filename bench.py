#!/usr/bin/env python
"""bench.py -- the FlatSparray hot path on B200 (one process per GPU).

A "step" is one pass of the hot path over one batch of synthetic input: BASELINE.json
config[1] -- two random (256,256,256,64) float32 operands at 1% density (~10.7 M nnz
each, 256 MB of operands: larger than L2, so nothing is re-used between steps) --
``c = a * b`` (intersection merge, flat.py:424-434) followed by ``c.sum(axis=2)``
(flat.py:607-626).  metric = binop-merge throughput in input nnz per second.

  value      : whole-job Gnnz/s with operands resident in HBM (CUDA events, max over ranks); steps are
               enqueued back to back -- results keep their element counts on the device (lazily
               counted FlatSparray results), the last step's result is resolved inside the timed region
  e2e        : same metric through the public FlatSparray API starting from pinned HOST
               buffers (H2D of all four operand arrays and D2H of the result inside the
               timed region)
  roofline   : the dominant kernel (fused merge-path intersection) timed alone with CUDA
               events on its launch stream; algorithmic bytes / time vs the measured HBM peak
  cpu_baseline / --impl reference : the reference's CPU path (its own Cython merge kernel
               from oracle/_ref when built, else the C port, + the numpy glue of flat.py
               restated in oracle/) timed on this box's host cores (single-threaded: the
               reference has no threading)

N > 1: weak scaling through ``sparray_b200.sharded.ShardedFlatSparray``.  ONE global array of shape
(256*N, 256, 256, 64) is range-partitioned on its flat index by SAMPLED splitters (quantiles of an
all-gathered sample of the stored indices, snapped to whole output rows of the axis-2 sum); both operands
share the splitters, so ``a*b`` and the axis-2 sum need no collective and ranks
only meet at the barriers.  Beside the headline every N (1 included) also runs BASELINE config[4] at its
stated per-GPU size -- the co-partitioned union ``a+b`` of two (128*N, 1024, 1024, 256) float32 arrays with
250 M stored elements per operand per GPU (2 G per operand at N=8), device-generated, sampled splitters --
and, for N > 1, the same with ``b`` misaligned (NCCL repartition and the fused NVLink peer-load merge):
``config4_union`` in the JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

SHAPE = (256, 256, 256, 64)
DENSITY = 0.01
DTYPE = np.float32
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# ncu --set full capture of this workload (profiles/README.md); None until a capture exists
NCU_TRAFFIC_BYTES = 210896384   # partition 6.88 MB + merge 195.89 MB read, 4.39 MB written + reorder 3.73 MB
NCU_TRAFFIC_SOURCE = ("profiles/r2_setops_ncu_raw.csv: dram__bytes_read.sum + dram__bytes_write.sum of the three launches of one "
                      "spb_intersect_binop call (merge_partition_kernel, setop_kernel<float,float,INTERSECT>, merge_reorder_kernel) "
                      "at this workload, ncu --set full")


def bench_config(na, nb):
    """The `config` object of the JSON line -- the SAME dict in both arms (ours / --impl reference)."""
    return {"workload": "BASELINE config[1]: (256,256,256,64) float32, 1% density, c=a*b (intersection) then "
                        "c.sum(axis=2); one such range shard per GPU (weak scaling)",
            "shape": list(SHAPE), "density": DENSITY, "value_dtype": "float32", "operand_seeds": [1, 2],
            "nnz_a": int(na), "nnz_b": int(nb)}


def load_oracle():
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import flat_oracle as fo
    import build_ref
    ref = build_ref.load()          # the reference's own Cython kernel, if it was built
    return fo, ref


def hbm_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def mark(self):
        """The timed region starts here: samples taken earlier (warm-up) are not reported, except
        the last one, which covers a region shorter than the 20 ms sampling period."""
        self.first = max(len(self.lines) - 1, 0)

    def wait_ready(self, timeout=3.0):
        """nvidia-smi's start-up talks to the driver and slows launches down: let it finish first."""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.02)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines[self.first:]:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def gen_operands(rank):
    """SURVEY.md section 8(d) generator; seeds a=1, b=2 (offset per rank for the sharded runs)."""
    fo, _ = load_oracle()
    ai, av = fo.random_operand(SHAPE, DENSITY, DTYPE, 1 + 100 * rank)
    bi, bv = fo.random_operand(SHAPE, DENSITY, DTYPE, 2 + 100 * rank)
    return ai, av, bi, bv


def cpu_step(fo, ref, ai, av, bi, bv):
    """The reference's CPU path for one step: flat.py:424-434 then flat.py:607-626."""
    if ref is not None:
        ci, la, lb = ref.intersect1d_sorted(ai, bi, return_inds=True)      # the reference's Cython kernel
        ci, cv = np.asarray(ci), np.multiply(av[la], bv[lb])
    else:
        ci, cv = fo.pairwise_intersect(ai, av, bi, bv, np.multiply)        # C port
    return fo.sum_axis(ci, cv, SHAPE, 2)


def run_reference(args, rank):
    if rank != 0:
        return
    fo, ref = load_oracle()
    ai, av, bi, bv = gen_operands(0)
    for _ in range(args.warmup):
        cpu_step(fo, ref, ai, av, bi, bv)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(fo, ref, ai, av, bi, bv)
    dt = (time.perf_counter() - t0) / args.steps
    val = (len(ai) + len(bi)) / dt / 1e9
    kind = "reference" if ref is not None else "port"
    sample = ("full config[1] step (%d + %d nnz) per step; merge kernel = %s; numpy glue restated in oracle/flat_oracle.py"
              % (len(ai), len(bi), "the reference's _merge.pyx compiled into oracle/_ref" if ref is not None
                 else "C port oracle/merge_oracle.c"))
    print(json.dumps({
        "impl": "reference", "metric": "FlatSparray binop merge throughput (a*b then sum(axis=2))",
        "value": val, "unit": "Gnnz/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(len(ai), len(bi)),
        "cpu_baseline": {"value": val, "unit": "Gnnz/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "Gnnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 500 (ours), 20 (--impl reference)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-ops", action="store_true", help="skip the per-op breakdown (other configs)")
    ap.add_argument("--eager", action="store_true", help="time plain API calls (never replay the captured step)")
    ap.add_argument("--graph", action="store_true", help="always replay the step captured into a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.steps is None:
        args.steps = 500 if args.impl == "ours" else 20

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: sparray_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # one process per GPU share the host cores: give every rank its own slice so that the Python that
        # enqueues the steps is not migrated or preempted by another rank's (the step is ~60 us of host
        # work against ~80 us of device work, so host jitter shows up in the max over ranks)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(len(cores) // world, 1)
            mine = cores[(local_rank * per) % len(cores):][:per]
            if mine:
                os.sched_setaffinity(0, set(mine))
        except (AttributeError, OSError):
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    import sparray_b200 as sp
    from sparray_b200 import _dtypes as dt
    from sparray_b200._lib import check, lib, ptr, stream, workspace

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ai, av, bi, bv = gen_operands(rank)
    na, nb = len(ai), len(bi)
    host = [torch.from_numpy(x).pin_memory() for x in (ai, av, bi, bv)]
    A = sp.FlatSparray(host[0].cuda(), host[1].cuda(), shape=SHAPE, is_canonical=True)
    B = sp.FlatSparray(host[2].cuda(), host[3].cuda(), shape=SHAPE, is_canonical=True)
    sharded_note = None
    if world > 1:
        # ONE global (256*N, 256, 256, 64) array: rank r generated block r; the shards are then re-cut at
        # SAMPLED splitters (snapped to whole output rows of the axis-2 sum) shared by both operands --
        # untimed setup, it goes through the NCCL exchange of ShardedFlatSparray.repartition
        from sparray_b200 import sharded as sh
        size = int(np.prod(SHAPE))
        gshape = (SHAPE[0] * world,) + tuple(SHAPE[1:])
        even = [r * size for r in range(world + 1)]
        a0 = sh.ShardedFlatSparray(A.indices + rank * size, A.data, gshape, even)
        b0 = sh.ShardedFlatSparray(B.indices + rank * size, B.data, gshape, even)
        both = torch.cat([a0.indices, b0.indices]).sort().values
        cuts = sh.sampled_splitters(both, gshape, samples_per_rank=4096, align=SHAPE[2] * SHAPE[3])
        A_sh, B_sh = a0.repartition(cuts), b0.repartition(cuts)
        del a0, b0, both
        na, nb = A_sh.nnz_local, B_sh.nnz_local           # what this rank merges per step
        sharded_note = {"class": "ShardedFlatSparray", "global_shape": list(gshape), "splitters": "sampled (4096 "
                        "samples per rank of the merged index stream, all-gathered; snapped down to multiples of "
                        "%d = one output row of sum(axis=2))" % (SHAPE[2] * SHAPE[3]),
                        "nnz_local_a_b": [na, nb], "collectives_in_step": 0}

        def step():
            c = A_sh * B_sh
            return c, c.sum(axis=2)

        def resolve(c, s):
            return s.nnz_local
    else:
        def step():
            c = A * B
            return c, c.sum(axis=2)

        def resolve(c, s):
            return s.nnz

    # ---- value: operands resident in HBM -------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    # Steps are enqueued asynchronously (no count read-back), so the host runs ahead of the device and
    # the caching allocator holds a few more blocks than in a synchronous loop: let it reach that
    # state before the warm-up proper, so no cudaMalloc lands in the timed region.
    prime = min(args.steps, 100)
    for _ in range(prime):
        c, s = step()
    resolve(c, s)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        c, s = step()
    resolve(c, s)
    # ---- the step as a CUDA graph: captured ONCE from the same public-API calls (the library is
    # capture-safe: a captured launch wipes the look-back words it uses with a memset node of its own), then
    # one cudaGraphLaunch per step instead of ~60 us of Python.  --eager times the plain calls instead.
    graph, launches_per_step, submission_probe = None, None, None
    if not args.eager:
        cap = torch.cuda.Stream()
        with torch.cuda.stream(cap):
            for _ in range(3):                       # scratch of the capture stream's workspace gets its size
                c, s = step()
            resolve(c, s)
        cap.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = lib.spb_launch_count()
        with torch.cuda.graph(graph, stream=cap):
            c, s = step()
        launches_per_step = int(lib.spb_launch_count() - l0)
        for _ in range(args.warmup):
            graph.replay()
        torch.cuda.synchronize()
        if not args.graph:
            # auto: whichever submission is faster HERE (max over ranks) -- plain calls keep their programmatic
            # dependent launches and win on an idle host; replays win when N ranks' interpreters share the cores
            def probe(fn, reps=50):
                barrier()
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                for _ in range(reps):
                    r = fn()
                p1.record()
                torch.cuda.synchronize()
                return max_over_ranks(p0.elapsed_time(p1) / reps)
            t_eager, t_graph = probe(step), probe(graph.replay)
            submission_probe = {"eager_ms": t_eager, "graph_ms": t_graph}
            if t_eager <= t_graph:
                graph = None
    # One untimed rehearsal of what the timed region ends with: the count read-back and the trimmed copies of the
    # last step's result.  torch.cuda.graph() empties the allocator's cache before it captures, so without this the
    # 1-2 MB blocks of those copies would be cudaMalloc'ed inside the timed region (~0.3 ms: nothing next to 500
    # steps, a fifth of a 20-step run).  The rehearsal resolves a throw-away wrapper around the same worst-case
    # buffers; `s` itself stays lazily counted.
    if graph is not None:
        graph.replay()
    else:
        c, s = step()
    torch.cuda.synchronize()
    pend = getattr(s, "_pending", None)
    try:                                   # (a rehearsal: whatever goes wrong here must not take the line down)
        if pend is not None:
            if world == 1:
                resolve(None, sp.FlatSparray._from_kernel(pend[0], pend[1], pend[2], s.shape))
            else:
                resolve(None, s._like(pend[0], pend[1], cnt=pend[2]))
    except Exception as exc:               # noqa: BLE001
        sys.stderr.write("bench: resolve rehearsal skipped (%r)\n" % (exc,))
    del pend
    torch.cuda.synchronize()
    barrier()
    sampler.mark()
    launches0 = lib.spb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    trace = [] if os.environ.get("SPB_BENCH_TRACE") else None      # diagnosis only: per-step events + host times
    for _ in range(args.steps):
        if trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            trace.append((ev, time.perf_counter()))
        if graph is not None:
            graph.replay()
        else:
            c, s = step()
    if trace is not None:
        trace.append((e1, time.perf_counter()))
    resolve(c, s)  # element counts stay on the device between steps; the last result is resolved in the timed region
    e1.record()
    barrier()
    launches = lib.spb_launch_count() - launches0
    if graph is not None:
        launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    if trace is not None:
        dev = [trace[i][0].elapsed_time(trace[i + 1][0]) for i in range(args.steps)]
        hst = [(trace[i + 1][1] - trace[i][1]) * 1e3 for i in range(args.steps)]
        top = sorted(range(args.steps), key=lambda i: -dev[i])[:5]
        sys.stderr.write("trace: total %.3f ms; slowest device steps %s; host %s\n" % (
            e0.elapsed_time(e1), [(i, round(dev[i], 3)) for i in top], [(i, round(hst[i], 3)) for i in top]))
    n_in = na + nb
    if world > 1:
        t = torch.tensor([n_in], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        n_in_total = int(t.item())
    else:
        n_in_total = n_in
    value = n_in_total / (ms_step * 1e-3) / 1e9
    nout_step = c.nnz if world == 1 else c.nnz_local
    na, nb = A.nnz, B.nnz                 # from here on: this rank's own block again (e2e, roofline, side lines)

    # ---- e2e: pinned host buffers -> public API -> host result -----------------------------------
    def e2e_step():
        a = sp.FlatSparray(host[0].cuda(non_blocking=True), host[1].cuda(non_blocking=True), shape=SHAPE,
                           is_canonical=True)
        b = sp.FlatSparray(host[2].cuda(non_blocking=True), host[3].cuda(non_blocking=True), shape=SHAPE,
                           is_canonical=True)
        r = (a * b).sum(axis=2)
        return r.indices.cpu(), r.data.cpu()
    for _ in range(3):
        ri, rd = e2e_step()
    e2e_steps = max(3, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ri, rd = e2e_step()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / e2e_steps * 1e3)
    barrier()
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = ri.numel() * 8 + rd.numel() * 8
    e2e_value = n_in_total / (e2e_ms * 1e-3) / 1e9

    # ---- roofline: the fused intersection kernel alone, CUDA events on its launch stream -----------
    cap = min(na, nb)
    out_idx = torch.empty(cap, dtype=torch.int64, device="cuda")
    out_val = torch.empty(cap, dtype=torch.float32, device="cuda")
    cnt = torch.empty(1, dtype=torch.int64, device="cuda")

    def launch_intersect():
        check(lib.spb_intersect_binop(dt.BINOPS["mul"], dt.SPB_F32, ptr(A.indices), ptr(A.data), na, ptr(B.indices),
                                      ptr(B.data), nb, ptr(out_idx), ptr(out_val), ptr(cnt), workspace(), stream()))
    for _ in range(5):
        launch_intersect()
    reps = 50
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a_, b_ in evs:
        a_.record()
        launch_intersect()
        b_.record()
    torch.cuda.synchronize()
    k_ms = float(np.mean([a_.elapsed_time(b_) for a_, b_ in evs]))
    nout = int(cnt.item())
    alg_bytes = (na + nb) * 8 + nout * (4 + 4) + nout * (8 + 4)
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "spb_intersect_binop = merge_partition + setop_kernel<float,float,INTERSECT> + "
                                          "merge_reorder (3 launches)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC_BYTES, "traffic_source": NCU_TRAFFIC_SOURCE,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "peak_source": peak_src,
                "timing": "CUDA events around each of %d launches on the launch stream; 256 MB of operands > 126 MB L2" % reps}

    # ---- per-op breakdown at the other BASELINE configs (N=1 only; not the headline) ---------------
    ops = None
    if world == 1 and not args.no_ops:
        ops = op_breakdown(sp, torch, A, B, peak)

    # ---- sharded exchange step (N>1 only; not the headline): operands with MISALIGNED splitters ------
    sharded = None
    if world > 1 and not args.no_ops:
        try:
            sharded = sharded_exchange(torch, dist, A, B, rank, world, barrier, max_over_ranks)
        except Exception as exc:  # never let the side measurement take the headline line down
            sharded = {"error": str(exc).splitlines()[0][:300]}

    # ---- BASELINE config[4] at its stated per-GPU size: sharded union a+b, 250 M nnz per operand per GPU -----
    config4 = None
    if not args.no_ops:
        del A, B
        torch.cuda.empty_cache()
        try:
            config4 = config4_union(sp, torch, dist, rank, world, barrier, max_over_ranks, peak)
        except Exception as exc:  # never let the side measurement take the headline line down
            config4 = {"error": str(exc).splitlines()[0][:300]}

    # ---- CPU baseline beside it (rank 0, N=1) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1:
        fo, ref = load_oracle()
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            cpu_step(fo, ref, ai, av, bi, bv)
            best = min(best, time.perf_counter() - t0)
        cpu = {"value": (na + nb) / best / 1e9, "unit": "Gnnz/s", "cores": 1,
               "kind": "reference" if ref is not None else "port", "ms_per_step": best * 1e3,
               "host_cores_available": os.cpu_count(),
               "sample": "the full step (whole workload, best of 3): %s merge kernel + numpy glue of flat.py restated in "
                         "oracle/flat_oracle.py; single-threaded like the reference"
                         % ("the reference's Cython (oracle/_ref)" if ref is not None else "C-port")}

    if rank == 0:
        out = {
            "metric": "FlatSparray binop merge throughput (a*b then sum(axis=2))",
            "value": value, "unit": "Gnnz/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(na, nb),
            "notes": {"nnz_out_of_the_product": nout_step, "input_nnz_all_ranks": n_in_total,
                      "per_gpu_shard": "one config[1]-sized range shard per GPU, co-partitioned (no collective)",
                      "l2": "operands (256 MB per GPU) exceed the 126 MB L2; no flush needed",
                      "sync": "steps are enqueued back to back: element counts stay on the device (lazily counted "
                              "results); the last step's result is resolved (count read back, buffers trimmed) "
                              "inside the timed region",
                      "step_submission": "eager API calls" if graph is None else
                                         "the step's public-API calls captured once into a CUDA graph (torch.cuda.graph), "
                                         "one replay per step; %d kernel launches per replay" % launches_per_step,
                      "submission_probe_ms": submission_probe,
                      "untimed_setup_steps": prime, "sharded": sharded_note},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Gnnz/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if ops is not None:
            out["ops"] = ops
        if sharded is not None:
            out["sharded_exchange"] = sharded
        if config4 is not None:
            out["config4_union"] = config4
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def time_op(torch, fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def synth_sorted(torch, n, size, dtype, seed):
    """Sorted duplicate-free flat indices of ~uniform density generated on the device (for
    configs too large to build on the host): cumulative sum of random positive gaps."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    gap = max(int(size // n), 1)
    steps = torch.randint(1, 2 * gap, (n,), device="cuda", generator=g, dtype=torch.int64) if gap > 1 else \
        torch.ones(n, device="cuda", dtype=torch.int64)
    idx = torch.cumsum(steps, 0) - 1
    idx = idx[idx < size]
    val = torch.randn(idx.numel(), device="cuda", generator=g, dtype=dtype)
    return idx.contiguous(), val


def sharded_exchange(torch, dist, A, B, rank, world, barrier, max_over_ranks):
    """SURVEY.md section 8(e), the one real exchange step of a binop: b arrives range-partitioned with
    splitters shifted by half a shard, so `a * b` first repartitions b -- counts all-gather + one
    variable-size all-to-all over NVLink in which every rank ships half of its shard to a neighbour --
    and then runs the local fused intersection.  Device-timed, max over ranks."""
    from sparray_b200 import sharded as sh
    size = int(np.prod(SHAPE))
    gshape = (SHAPE[0] * world,) + tuple(SHAPE[1:])
    even = [r * size for r in range(world + 1)]
    shifted = [0] + [r * size + size // 2 for r in range(world - 1)] + [world * size]     # rank 0 owns half a shard
    a_sh = sh.ShardedFlatSparray(A.indices + rank * size, A.data, gshape, even)
    # build b on the shifted partition: take it from the even one with a (untimed) repartition
    b_even = sh.ShardedFlatSparray(B.indices + rank * size, B.data, gshape, even)
    b_sh = b_even.repartition(shifted)
    moved = b_sh.repartition(even)                  # warm-up of the timed exchange (allocator, NCCL channels)
    n_moved_local = int(b_sh.nnz_local)
    del moved

    def timed(fn, reps=5):
        ts = []
        for _ in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(max_over_ranks(e0.elapsed_time(e1)))
            del r
        return float(np.median(ts))
    t_rep = timed(lambda: b_sh.repartition(even))
    t_mul = timed(lambda: (a_sh * b_sh).nnz_local)          # repartition + local fused intersection
    # the same product with no exchange step: b's shards published in symmetric memory (untimed, once per
    # operand), the merge kernels read the partner's slices over NVLink inside their TMA staging
    t_peer, peer_note = None, "ok"
    t_pull = t_pull_mul = None
    try:
        b_pub = b_sh.publish()
        want = (a_sh * b_sh)
        got = a_sh.binop_peer(b_pub, "intersect", "mul")
        same = got.nnz_local == want.nnz_local and bool((got.indices == want.indices).all()) and \
            bool((got.data == want.data).all())
        del want, got
        if same:
            t_peer = timed(lambda: a_sh.binop_peer(b_pub, "intersect", "mul").nnz_local)
        else:
            peer_note = "MISMATCH against the repartition path"
        # the repartition itself, PULLED from the published shards (peer copies over NVLink, no NCCL send / recv)
        pulled = b_pub.repartition(even, b_sh.ops)
        ref = b_sh.repartition(even)
        if bool((pulled.indices == ref.indices).all()) and bool((pulled.data == ref.data).all()):
            t_pull = timed(lambda: b_pub.repartition(even, b_sh.ops))
            t_pull_mul = timed(lambda: (a_sh * b_pub.repartition(even, b_sh.ops)).nnz_local)
        else:
            peer_note = "MISMATCH of the pulled repartition"
        del pulled, ref
        b_pub.barrier()
    except Exception as exc:      # symmetric memory unavailable on this box: the NCCL path above still stands
        peer_note = "unavailable: %s" % str(exc).splitlines()[0][:200]
    # bytes a rank SENDS in the exchange: the elements of its shifted shard that belong to another rank
    sent = int(((b_sh.indices < rank * size) | (b_sh.indices >= (rank + 1) * size)).sum().item())
    tot = torch.tensor([sent, n_moved_local], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    sent_all, n_all = int(tot[0].item()), int(tot[1].item())
    bytes_per_elem = 8 + B.data.element_size()
    return {"what": "b range-partitioned with splitters shifted by half a shard; a*b = repartition(b) + local intersection",
            "repartition_ms": t_rep, "mul_with_repartition_ms": t_mul, "mul_with_peer_loads_ms": t_peer, "peer_loads_status": peer_note,
            "elements_repartitioned": n_all,
            "elements_crossing_gpus": sent_all,
            "nvlink_GBps_per_gpu_send": sent_all / world * bytes_per_elem / (t_rep * 1e-3) / 1e9,
            "repartition_pulled_ms": t_pull, "mul_with_pulled_repartition_ms": t_pull_mul,
            "nvlink_GBps_per_gpu_pulled": (sent_all / world * bytes_per_elem / (t_pull * 1e-3) / 1e9) if t_pull else None,
            "nvlink_peak_GBps_per_direction": 900.0,
            "note": "repartition_pulled = PeerShards.repartition: every rank copies its slices straight out of its peers' "
                    "published shards (symmetric memory, device copies over NVLink; one all-gather of cut positions).  "
                    "repartition = lower_bound of the splitters, counts all-gather (host sync), index and value "
                    "all_to_all_single (NCCL); received segments are already in order, no merge.  peer loads = "
                    "ShardedFlatSparray.binop_peer: one all-gather of cut positions, then one fused merge launch "
                    "per overlapping source rank reading that rank's slice from symmetric memory over NVLink"}


def config4_union(sp, torch, dist, rank, world, barrier, max_over_ranks, peak):
    """BASELINE config[4]: union a+b of two 4-D float32 arrays with 250 M stored
    elements per operand PER GPU -- global shape (128*N, 1024, 1024, 256), 2 G per operand at N=8 -- generated
    on the device, range-partitioned by sampled splitters, co-partitioned (no collective); then, N > 1, with b
    cut at splitters shifted by half a shard: NCCL repartition and the fused peer-load merge."""
    from sparray_b200 import sharded as sh
    slab = (128, 1024, 1024, 256)
    size = int(np.prod(slab))
    n_per = 250_000_000
    gshape = (slab[0] * world,) + slab[1:]
    ai, av = synth_sorted(torch, n_per, size, torch.float32, 11 + 100 * rank)
    bi, bv = synth_sorted(torch, n_per, size, torch.float32, 12 + 100 * rank)

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            r = fn()
            del r
        ts = []
        for _ in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(max_over_ranks(e0.elapsed_time(e1)))
            del r
        return float(np.median(ts))

    if world == 1:
        A = sp.FlatSparray(ai, av, shape=gshape, is_canonical=True)
        B = sp.FlatSparray(bi, bv, shape=gshape, is_canonical=True)
        na, nb = A.nnz, B.nnz
        nout = (A + B).nnz
        ms = timed(lambda: A + B)
        tot_in, tot_out = na + nb, nout
        extra = {}
    else:
        even = [r * size for r in range(world + 1)]
        ai += rank * size
        bi += rank * size
        a0 = sh.ShardedFlatSparray(ai, av, gshape, even)
        b0 = sh.ShardedFlatSparray(bi, bv, gshape, even)
        cuts = sh.sampled_splitters(a0.indices, gshape, samples_per_rank=4096)
        A, B = a0.repartition(cuts), b0.repartition(cuts)
        del a0, b0, ai, av, bi, bv
        torch.cuda.empty_cache()
        na, nb = A.nnz_local, B.nnz_local
        nout = (A + B).nnz_local
        ms = timed(lambda: A + B)
        t = torch.tensor([na + nb, nout], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        tot_in, tot_out = int(t[0].item()), int(t[1].item())
        # b misaligned: its splitters sit half a shard to the right of a's
        shifted = [cuts[0]] + [(cuts[r] + cuts[r + 1]) // 2 for r in range(world - 1)] + [cuts[-1]]
        B_sh = B.repartition(shifted)
        moved = int(((B_sh.indices < cuts[rank]) | (B_sh.indices >= cuts[rank + 1])).sum().item())
        t_rep = timed(lambda: B_sh.repartition(cuts), reps=3, warm=1)
        t_add = timed(lambda: (A + B_sh).nnz_local, reps=3, warm=1)
        mv = torch.tensor([moved], dtype=torch.int64, device="cuda")
        dist.all_reduce(mv)
        extra = {"misaligned_b": {"what": "b cut at splitters shifted by half a shard: a+b = repartition(b) + local union",
                                  "repartition_ms": t_rep, "add_with_repartition_ms": t_add,
                                  "elements_crossing_gpus": int(mv.item()),
                                  "nvlink_GBps_per_gpu_send": int(mv.item()) / world * 12 / (t_rep * 1e-3) / 1e9,
                                  "nvlink_peak_GBps_per_direction": 900.0, "nvlink_measured_peer_copy_GBps": 770.0}}
        try:
            b_pub = B_sh.publish()
            t_peer = timed(lambda: A.binop_peer(b_pub, "union", "add").nnz_local, reps=3, warm=1)
            extra["misaligned_b"]["add_with_peer_loads_ms"] = t_peer
            t_pull = timed(lambda: b_pub.repartition(cuts, B_sh.ops), reps=3, warm=1)
            t_pull_add = timed(lambda: (A + b_pub.repartition(cuts, B_sh.ops)).nnz_local, reps=3, warm=1)
            extra["misaligned_b"]["repartition_pulled_ms"] = t_pull
            extra["misaligned_b"]["add_with_pulled_repartition_ms"] = t_pull_add
            extra["misaligned_b"]["nvlink_GBps_per_gpu_pulled"] = int(mv.item()) / world * 12 / (t_pull * 1e-3) / 1e9
            b_pub.barrier()
        except Exception as exc:
            extra["misaligned_b"]["peer_loads"] = "unavailable: %s" % str(exc).splitlines()[0][:200]
    alg_bytes_per_gpu = (na + nb + nout) * 12
    out = {"workload": "BASELINE config[4]: union a+b of two (128*N,1024,1024,256) float32 arrays, 250 M stored elements "
                       "per operand per GPU, sampled splitters, co-partitioned (no collective), device-generated",
           "global_shape": list(gshape), "nnz_in_all_gpus": tot_in, "nnz_out_all_gpus": tot_out,
           "ms": ms, "Gnnz_per_s": tot_in / (ms * 1e-3) / 1e9,
           "alg_GBps_per_gpu": alg_bytes_per_gpu / (ms * 1e-3) / 1e9,
           "frac_of_hbm_peak_per_gpu": alg_bytes_per_gpu / (ms * 1e-3) / 1e9 / peak,
           "timing": "CUDA events around the API call (lazily counted result), max over ranks, median of 5"}
    out.update(extra)
    return out


def op_breakdown(sp, torch, A, B, peak):
    """API-level timings at the other BASELINE configs (CUDA events around the call; binops and dense
    axis sums return lazily counted results, so their count read-back is not inside; the sort-based
    ops, last-axis sums and SpMV synchronise as part of the call)."""
    res = {}

    def rec(name, n_in, alg_bytes, fn, reps=10, warm=3):
        med, best = time_op(torch, fn, reps=reps, warm=warm)
        res[name] = {"ms": med, "ms_best": best, "Gnnz_per_s": n_in / (med * 1e-3) / 1e9,
                     "alg_GBps": alg_bytes / (med * 1e-3) / 1e9, "frac_of_hbm_peak": alg_bytes / (med * 1e-3) / 1e9 / peak}
    na, nb = A.nnz, B.nnz
    u = A + B
    rec("c2_union_a+b_f32", na + nb, (na + nb + u.nnz) * 12, lambda: A + B)
    c = A * B
    rec("c2_intersect_a*b_f32", na + nb, (na + nb) * 8 + c.nnz * 8 + c.nnz * 12, lambda: A * B)
    s = A.sum(axis=2)
    rec("c2_sum_axis2_full_operand", na, na * 12 + s.nnz * 16, lambda: A.sum(axis=2))
    s = A.sum(axis=3)
    rec("c2_sum_axis3_last_axis", na, na * 12 + s.nnz * 16, lambda: A.sum(axis=3))
    rec("c2_transpose_3210", na, 2 * na * 12, lambda: A.transpose(3, 2, 1, 0))
    del u, c, s
    # config[0]: 10000 x 10000 float64 at 1%
    n1 = 1_000_000
    i1, v1 = synth_sorted(torch, n1, 10 ** 8, torch.float64, 1)
    i2, v2 = synth_sorted(torch, n1, 10 ** 8, torch.float64, 2)
    X = sp.FlatSparray(i1, v1, shape=(10000, 10000), is_canonical=True)
    Y = sp.FlatSparray(i2, v2, shape=(10000, 10000), is_canonical=True)
    u = X + Y
    rec("c1_union_a+b_f64", X.nnz + Y.nnz, (X.nnz + Y.nnz + u.nnz) * 16, lambda: X + Y, reps=20)
    rec("c1_intersect_a*b_f64", X.nnz + Y.nnz, (X.nnz + Y.nnz) * 8, lambda: X * Y, reps=20)
    # config[2] literal: (4096,4096,1024) float32 at 1e-4 -> 1.7 M nnz, swapaxes(0,2)
    i3, v3 = synth_sorted(torch, 1_717_987, 2 ** 34, torch.float32, 3)
    Z = sp.FlatSparray(i3, v3, shape=(4096, 4096, 1024), is_canonical=True)
    rec("c3_literal_swapaxes02_f32", Z.nnz, 2 * Z.nnz * 12, lambda: Z.swapaxes(0, 2))
    # config[3]: 1M x 1M float64, 100 M nnz: SpMV and sum(axis=1)
    i4, v4 = synth_sorted(torch, 100_000_000, 10 ** 12, torch.float64, 4)
    W = sp.FlatSparray(i4, v4, shape=(10 ** 6, 10 ** 6), is_canonical=True)
    x = torch.randn(10 ** 6, device="cuda", dtype=torch.float64)
    from sparray_b200 import _kernels as k
    rec("c4_spmv_f64_100M", W.nnz, W.nnz * 16 + 2 * 8 * 10 ** 6, lambda: k.spmv(W.indices, W.data, 10 ** 6, 10 ** 6, x, torch.float64), reps=5)
    s = W.sum(axis=1)
    rec("c4_sum_axis1_f64_100M", W.nnz, W.nnz * 16 + s.nnz * 16, lambda: W.sum(axis=1), reps=5)
    rec("c4_transpose_f64_100M", W.nnz, 2 * W.nnz * 16, lambda: W.T, reps=3)
    del W, s, i4, v4, X, Y, Z, u
    torch.cuda.empty_cache()
    # config[2] at its stated size: (4096,4096,1024) float32 at density 0.1 -> 1.7 G nnz (20.6 GB), swapaxes(0,2)
    free = torch.cuda.mem_get_info()[0]
    if free > 100 * 2 ** 30:
        i5, v5 = synth_sorted(torch, 1_717_986_918, 2 ** 34, torch.float32, 5)
        Zb = sp.FlatSparray(i5, v5, shape=(4096, 4096, 1024), is_canonical=True)
        del i5, v5
        rec("c3_stated_size_swapaxes02_f32_1.7G", Zb.nnz, 2 * Zb.nnz * 12, lambda: Zb.swapaxes(0, 2), reps=3, warm=1)
        del Zb
        torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    main()
