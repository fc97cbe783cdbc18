"""Range-sharded FlatSparray across the GPUs of one NVSwitch box (one process per GPU,
``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).

The reference has no distributed code at all: this layer is new
design.  An array is split by contiguous FLAT-INDEX RANGES -- rank r owns the stored
elements with ``splitters[r] <= flat index < splitters[r + 1]`` -- so every shard is itself
a canonical (sorted, duplicate-free) ``(indices, data)`` pair and the single-GPU kernels
run on it unchanged.

  * co-partitioned binops (same splitters on both operands): purely local, NO collective;
  * misaligned operands: ``repartition`` = allgather of the send counts + ONE variable-size
    all-to-all of (index, value) segments.  Each shard is sorted, so the segment for every
    destination is a contiguous slice (found by a batched binary search of the splitters),
    and because the source partition is itself a range partition, concatenating the
    received segments in source-rank order is already globally sorted: no merge, no sort;
  * misaligned operands, second form (``publish`` + ``binop_peer``): the partner's shards live in
    symmetric memory (``torch.distributed._symmetric_memory``), every GPU maps every peer's copy, and
    the fused merge kernels READ THE PARTNER'S SLICES OVER NVLINK inside their TMA staging -- the
    exchange and the merge are one kernel, nothing is repartitioned, no intermediate buffer exists;
    the only collective is one tiny all-gather of cut positions;
  * axis sums: a range shard can only touch the output rows of its own flat range, so every rank reduces
    locally into a WINDOW of the reduced array (dense accumulators for that window only).  When the
    splitters do not cut through a block of the reduced axis (``sampled_splitters(..., align=...)``)
    that is the whole op: NO collective.  Otherwise the partial sums of the cut rows travel to the rank
    that owns the row (one variable-size exchange) and are merged with the union-add kernel.  When every
    rank touches every output cell (the leading axis is reduced) and the reduced array is small:
    dense per-GPU accumulators + all-reduce;
  * transpose: local re-key + sort (only the bits that matter), exchange of contiguous slices by new-key
    splitters, then a pairwise merge of the received sorted runs (union kernel) -- no second sort;
  * SpMV: rows follow flat ranges (row-major), x is replicated, y partials are all-reduced.

The kernels are reached through a small ``ops`` object so the orchestration (splitter
choice, counts exchange, all-to-all plumbing, border fix-up) can be exercised with
world_size 2 on CPU under gloo; the shipped default is ``CudaOps`` and there is no CPU
fallback -- the tests inject their own checker ops.
"""
import numpy as np
import torch
import torch.distributed as dist


class CudaOps:
    """Local kernels on this rank's GPU (the product path)."""

    def __init__(self):
        from . import _kernels
        self.k = _kernels

    def lower_bound(self, sorted_idx, query):
        return self.k.lower_bound(sorted_idx, query, want_found=False)[0]

    def binop(self, kind, op, a_idx, a_val, b_idx, b_val):
        from . import _dtypes as dt
        if a_val.dtype != b_val.dtype:
            common = dt.promote(a_val.dtype, b_val.dtype)
            a_val = self.k.cast(a_val, common) if a_val.dtype != common else a_val
            b_val = self.k.cast(b_val, common) if b_val.dtype != common else b_val
        return self.k.setop_binop(kind, op, a_idx, a_val, b_idx, b_val)[:2]

    def binop_lazy(self, kind, op, a_idx, a_val, b_idx, b_val):
        """(worst-case-sized indices, values, device count or None): no host synchronisation."""
        from . import _dtypes as dt
        if a_val.dtype != b_val.dtype:
            common = dt.promote(a_val.dtype, b_val.dtype)
            a_val = self.k.cast(a_val, common) if a_val.dtype != common else a_val
            b_val = self.k.cast(b_val, common) if b_val.dtype != common else b_val
        return self.k.setop_binop(kind, op, a_idx, a_val, b_idx, b_val, lazy=True)

    def reduce_by_key(self, keys, vals, divisor=1):
        return self.k.reduce_by_key(keys, vals, divisor)

    def sum_axis_window(self, idx, val, shape, axis, cell_lo, cell_hi, n_dev=None):
        """Local sum(axis) of a shard whose output cells all lie in [cell_lo, cell_hi):
        (sorted unique keys, float64 sums, device count or None)."""
        return self.k.sum_axis(idx, val, shape, axis, n_dev=n_dev, window=(cell_lo, cell_hi))

    def can_chain_count(self, cells):
        """True when sum_axis_window on a window of `cells` cells takes the producer's count from the device."""
        return 0 < cells <= self.k._l2_resident_cells(self.k._cur_dev())

    def rekey_sort(self, idx, val, in_s, out_s, div, bits):
        return self.k.rekey_sort(idx, val, in_s, out_s, div, bits)

    def axis_accumulate(self, idx, val, in_s, out_s, cells):
        return self.k.axis_accumulate(idx, val, in_s, out_s, cells)

    def emit_touched(self, acc, touched, lo, hi):
        return self.k.emit_touched(acc, touched, lo, hi)

    def spmv(self, idx, val, ncols, nrows, x):
        return self.k.spmv(idx, val, ncols, nrows, x, val.dtype)

    def device(self):
        return torch.device("cuda", torch.cuda.current_device())


DENSE_ALLREDUCE_MAX_CELLS = 1 << 24      # 128 MiB of float64 accumulators per GPU


def _size(shape):
    return int(np.prod([int(s) for s in shape], dtype=object))


def even_splitters(shape, world):
    """world + 1 flat-index boundaries of (almost) equal width, aligned to whole rows of the
    leading axis when that is possible (so that last-axis reductions and SpMV never straddle)."""
    size = _size(shape)
    row = size // int(shape[0]) if shape and shape[0] else 1
    nrows = size // row if row else 0
    if nrows >= world:
        cuts = [(nrows * r // world) * row for r in range(world)]
    else:
        cuts = [size * r // world for r in range(world)]
    return cuts + [size]


def sampled_splitters(local_idx, shape, group=None, ops=None, samples_per_rank=256, align=1):
    """Splitters that balance nnz: every rank contributes evenly spaced samples of its sorted
    indices, the samples are all-gathered (a few KB), and the r/world quantiles become the
    boundaries.  align > 1 snaps every interior boundary down to a multiple
    of `align` flat indices -- e.g. ``prod(shape[axis:])`` so that ``sum(axis)`` never cuts an
    output row and needs no exchange."""
    world = dist.get_world_size(group)
    n = local_idx.numel()
    take = min(samples_per_rank, n)
    if take:
        # evenly spaced positions in exact integer arithmetic (a float32 linspace overshoots n - 1 beyond 2^24)
        pos = (torch.arange(take, device=local_idx.device, dtype=torch.int64) * (n - 1)) // max(take - 1, 1)
        sample = local_idx[pos]
    else:
        sample = local_idx[:0]
    padded = torch.full((samples_per_rank,), -1, dtype=torch.int64, device=local_idx.device)
    padded[:take] = sample
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    allv = torch.cat(gathered)
    allv = allv[allv >= 0].cpu().numpy()
    allv.sort()
    size = _size(shape)
    if len(allv) == 0:
        return even_splitters(shape, world)
    align = max(int(align), 1)
    cuts = [0] + [int(allv[len(allv) * r // world]) // align * align for r in range(1, world)] + [size]
    for r in range(1, world + 1):          # keep them monotone
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts


class ShardedFlatSparray:
    """This rank's shard of a range-partitioned FlatSparray.  Like FlatSparray, results whose size
    depends on the data are lazily counted: the worst-case-sized buffers and the device-side count
    are kept until ``indices`` / ``data`` / ``nnz_local`` are looked at, so a chain such as
    ``(a * b).sum(axis)`` on co-partitioned, row-aligned shards never waits for the host."""

    def __init__(self, indices, data, shape, splitters, group=None, ops=None, _pending=None, _like=None):
        self._indices, self._data, self._pending = indices, data, _pending
        self.group = group
        if _like is not None:
            # derived from another shard of the same group: skip the (slow) process-group queries and the
            # re-validation of shape / splitters that the parent already did
            self.shape, self.splitters = shape, splitters
            self.ops, self.rank, self.world = _like.ops, _like.rank, _like.world
            return
        self.shape = tuple(int(s) for s in shape)
        self.splitters = [int(s) for s in splitters]
        self.ops = ops if ops is not None else CudaOps()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        assert len(self.splitters) == self.world + 1

    def _resolve(self):
        idx, val, cnt = self._pending
        n = int(cnt.item())
        self._indices, self._data, self._pending = idx[:n], val[:n], None

    @property
    def indices(self):
        if self._pending is not None:
            self._resolve()
        return self._indices

    @property
    def data(self):
        if self._pending is not None:
            self._resolve()
        return self._data

    @property
    def nnz_local(self):
        return int(self.indices.numel())

    def nnz(self):
        t = torch.tensor([self.nnz_local], dtype=torch.int64, device=self.indices.device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def _like(self, indices, data, shape=None, splitters=None, cnt=None):
        shape = self.shape if shape is None else tuple(int(s) for s in shape)
        splitters = self.splitters if splitters is None else [int(s) for s in splitters]
        if cnt is not None:
            return ShardedFlatSparray(None, None, shape, splitters, self.group, _pending=(indices, data, cnt), _like=self)
        return ShardedFlatSparray(indices, data, shape, splitters, self.group, _like=self)

    # ---- co-partitioned / repartitioned binops -------------------------------------------------------
    def _binop(self, other, kind, op):
        assert self.shape == other.shape
        if isinstance(other, PeerShards):
            # a published operand: its slices are pulled out of the peers' memory over NVLink (526 GB/s per GPU at
            # BASELINE config 5's size on 8 GPUs, against 236 GB/s for the NCCL send / receive group below)
            other = other.repartition(self.splitters, self.ops)
        elif other.splitters != self.splitters:
            other = other.repartition(self.splitters)       # the only communication of a binop
        lazy = getattr(self.ops, "binop_lazy", None)
        if lazy is not None:
            idx, val, cnt = lazy(kind, op, self.indices, self.data, other.indices, other.data)
            return self._like(idx, val, cnt=cnt)
        idx, val = self.ops.binop(kind, op, self.indices, self.data, other.indices, other.data)
        return self._like(idx, val)

    def __add__(self, other):
        return self._binop(other, "union", "add")

    def __sub__(self, other):
        return self._binop(other, "union", "sub")

    def __mul__(self, other):
        return self._binop(other, "intersect", "mul")

    def minimum(self, other):
        return self._binop(other, "union", "minimum")

    def maximum(self, other):
        return self._binop(other, "union", "maximum")

    # ---- misaligned operands without an exchange step: read the partner over NVLink ---------------------
    def publish(self):
        """Copy this shard into symmetric memory and map every peer's copy (collective; do it once for
        an operand that will be the misaligned partner of several ops)."""
        return PeerShards(self)

    def binop_peer(self, other, kind, op):
        """``self (op) other`` where ``other`` is a published, differently partitioned operand: for
        every source rank whose key range overlaps mine, ONE fused merge launch whose B operand is a
        slice of that rank's shard read straight from its memory (TMA bulk copies over NVLink).  The
        result is partitioned like ``self``.  Collectives: one all-gather of (world + 1) cut positions."""
        assert isinstance(other, PeerShards) and other.shape == self.shape
        dev = self.indices.device
        lo, hi = self.splitters[self.rank], self.splitters[self.rank + 1]
        # where my key range [lo, hi) starts and ends inside EVERY rank's shard of `other`: each rank
        # cuts its own shard at all my-partition splitters, one all-gather shares the cuts
        mine = torch.tensor(self.splitters, dtype=torch.int64, device=dev)
        cuts = self.ops.lower_bound(other.local_indices(), mine)                       # [world + 1]
        all_cuts = torch.empty(self.world * (self.world + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_cuts, cuts.contiguous(), group=self.group)
        # my own elements, cut at the partner's splitters (clipped to my range): piece s meets source s
        theirs = torch.tensor([min(max(x, lo), hi) for x in other.splitters], dtype=torch.int64, device=dev)
        a_cuts = self.ops.lower_bound(self.indices, theirs)
        host = torch.cat([all_cuts, a_cuts]).cpu().tolist()                           # the one read-back
        all_cuts, a_cuts = host[:self.world * (self.world + 1)], host[self.world * (self.world + 1):]
        # all launches first (their element counts stay on the device), then ONE read-back of the counts
        pieces = []
        for s in range(self.world):
            c0 = all_cuts[s * (self.world + 1) + self.rank]
            c1 = all_cuts[s * (self.world + 1) + self.rank + 1]
            a0, a1 = a_cuts[s], a_cuts[s + 1]
            if kind == "intersect" and (c1 == c0 or a1 == a0):
                continue
            if c1 == c0 and a1 == a0:
                continue
            b_idx, b_val = other.slice_of(s, c0, c1)
            pieces.append(self.ops.binop_lazy(kind, op, self.indices[a0:a1], self.data[a0:a1], b_idx, b_val))
        if not pieces:
            i, v = self.ops.binop(kind, op, self.indices[:0], self.data[:0], self.indices[:0], other.local_values()[:0])
            return self._like(i, v)
        pending = [p[2] for p in pieces if p[2] is not None]
        counts = torch.cat(pending).cpu().tolist() if pending else []
        out_i, out_v, j = [], [], 0
        for i, v, c in pieces:
            if c is not None:
                i, v = i[:counts[j]], v[:counts[j]]
                j += 1
            out_i.append(i)
            out_v.append(v)
        return self._like(torch.cat(out_i), torch.cat(out_v))

    # ---- repartition: one count exchange + one variable-size exchange ----------------------------------
    def repartition(self, new_splitters):
        new_splitters = [int(s) for s in new_splitters]
        recv_idx, recv_val, _ = _exchange_runs(self.indices, self.data, new_splitters, self.group, self.ops)
        # sources are range-ordered, so the concatenation in source order is globally sorted
        return self._like(recv_idx, recv_val, splitters=new_splitters)

    # ---- reductions -------------------------------------------------------------------------------------
    def sum(self, axis):
        """sum over one axis -> ShardedFlatSparray with float64 data (flat.py:607-626)."""
        axis = int(axis) % len(self.shape)
        new_shape = self.shape[:axis] + self.shape[axis + 1:]
        if not new_shape:
            new_shape = (1,)
        span = self.shape[axis]
        inner = _size(self.shape[axis + 1:])             # cells of one output row
        block = span * inner                             # flat indices that fold into one output row
        cells_total = _size(self.shape) // max(span, 1)
        sp = self.splitters
        # output rows each rank's flat range can touch: [row_lo, row_hi]
        rows = [(sp[r] // block, (sp[r + 1] - 1) // block) for r in range(self.world) if sp[r + 1] > sp[r]]
        window_cells = sum((hi - lo + 1) * inner for lo, hi in rows)
        if cells_total <= DENSE_ALLREDUCE_MAX_CELLS and window_cells >= 2 * cells_total:
            # every rank touches (nearly) every output cell -- the leading axis is reduced -- and the reduced
            # array is small enough to hold densely on every GPU: local float64
            # accumulators, one all-reduce of the sums and one of the "touched" flags, then every rank emits
            # its own range of output cells (ranges aligned to 64 cells)
            from . import _plan
            in_s, out_s = _plan.rekey_strides(self.shape, [a for a in range(len(self.shape)) if a != axis])
            acc, touched = self.ops.axis_accumulate(self.indices, self.data, in_s, out_s, cells_total)
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(touched, op=dist.ReduceOp.MAX, group=self.group)
            per = -(-cells_total // self.world)
            per = -(-per // 64) * 64
            out_split = [min(r * per, cells_total) for r in range(self.world)] + [cells_total]
            keys, sums = self.ops.emit_touched(acc, touched, out_split[self.rank], out_split[self.rank + 1])
            return ShardedFlatSparray(keys, sums, new_shape, out_split, self.group, self.ops)
        # local reduction into the window of output rows my flat range touches
        lo, hi = sp[self.rank], sp[self.rank + 1]
        row_lo = lo // block
        row_hi = (hi - 1) // block if hi > lo else row_lo - 1
        cell_lo, cell_hi = row_lo * inner, (row_hi + 1) * inner
        if self._pending is not None and self.ops.can_chain_count(cell_hi - cell_lo):
            p_idx, p_val, p_cnt = self._pending          # the producer's count never leaves the device
            keys, sums, cnt = self.ops.sum_axis_window(p_idx, p_val, self.shape, axis, cell_lo, cell_hi, n_dev=p_cnt)
        else:
            keys, sums, cnt = self.ops.sum_axis_window(self.indices, self.data, self.shape, axis, cell_lo, cell_hi)
        # an output row cut by a splitter belongs to the rank on the right of the cut
        new_split = [(s // block) * inner for s in sp[:-1]] + [cells_total]
        if all(s % block == 0 for s in sp[1:-1]):
            return self._like(keys, sums, shape=new_shape, splitters=new_split, cnt=cnt)
        if cnt is not None:
            n = int(cnt.item())
            keys, sums = keys[:n], sums[:n]
        # the partial sums of cut rows go to the row's owner: every destination receives sorted runs that
        # can only overlap inside its first row; merge them with the union-add kernel
        recv_idx, recv_val, counts = _exchange_runs(keys, sums, new_split, self.group, self.ops)
        keys, sums = _merge_runs(self.ops, recv_idx, recv_val, counts, "add")
        return ShardedFlatSparray(keys, sums, new_shape, new_split, self.group, self.ops)

    # ---- SpMV --------------------------------------------------------------------------------------------
    def dot(self, x):
        """a.dot(x) for a replicated dense vector x -> replicated dense y (float64 partials all-reduced)."""
        ncols = self.shape[-1]
        nrows = _size(self.shape) // ncols
        y = self.ops.spmv(self.indices, self.data, ncols, nrows, x)
        y = y.to(torch.float64)
        dist.all_reduce(y, group=self.group)
        return y.reshape(self.shape[:-1])

    # ---- transpose ---------------------------------------------------------------------------------------
    def transpose(self, *axes):
        """Local re-key + sort of the bits that matter, exchange of contiguous slices by new-key range,
        pairwise merge of the received sorted runs (their keys are disjoint) -- no second sort."""
        from . import _plan
        ndim = len(self.shape)
        axes = tuple(range(ndim - 1, -1, -1)) if not axes else tuple(int(a) for a in axes)
        new_shape = tuple(self.shape[a] for a in axes)
        in_s, out_s, div, bits = _plan.transpose_plan(self.shape, axes)
        keys, vals = self.ops.rekey_sort(self.indices, self.data, in_s, out_s, div, bits)
        out_split = even_splitters(new_shape, self.world)
        recv_idx, recv_val, counts = _exchange_runs(keys, vals, out_split, self.group, self.ops)
        k2, v2 = _merge_runs(self.ops, recv_idx, recv_val, counts, "override")
        return ShardedFlatSparray(k2, v2, new_shape, out_split, self.group, self.ops)

    # ---- gather to one canonical array (tests / small results) --------------------------------------------
    def gather(self):
        """All ranks get the full (indices, data) as host numpy arrays."""
        n = torch.tensor([self.nnz_local], dtype=torch.int64, device=self.indices.device)
        ns = [torch.empty_like(n) for _ in range(self.world)]
        dist.all_gather(ns, n, group=self.group)
        ns = [int(t.item()) for t in ns]
        cap = max(ns + [1])
        pi = torch.zeros(cap, dtype=torch.int64, device=self.indices.device)
        pv = torch.zeros(cap, dtype=self.data.dtype, device=self.indices.device)
        pi[:self.nnz_local] = self.indices
        pv[:self.nnz_local] = self.data
        gi = [torch.empty_like(pi) for _ in range(self.world)]
        gv = [torch.empty_like(pv) for _ in range(self.world)]
        dist.all_gather(gi, pi, group=self.group)
        dist.all_gather(gv, pv, group=self.group)
        idx = np.concatenate([gi[r][:ns[r]].cpu().numpy() for r in range(self.world)])
        val = np.concatenate([gv[r][:ns[r]].cpu().numpy() for r in range(self.world)])
        return idx, val


def _exchange_runs(keys, vals, splitters, group, ops):
    """Variable-size exchange of a SORTED local (keys, vals) array by destination key range: destination d
    gets my elements in [splitters[d], splitters[d + 1]) -- a contiguous slice, found by a batched binary
    search.  One small device-side all-gather of the cut positions, ONE host read-back (of the gathered
    matrix, which holds my own cuts too), then the index and the value slices of all peers travel in one
    batch of point-to-point operations (a single NCCL group).
    Returns (received keys, received values, per-source counts); the received array is the concatenation
    of one sorted run per source, in source order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = keys.device
    bounds = torch.tensor([int(s) for s in splitters], dtype=torch.int64, device=dev)
    cut = ops.lower_bound(keys, bounds).contiguous()                              # [world + 1]
    all_cut = torch.empty(world * (world + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_cut, cut, group=group)
    cuts = all_cut.cpu().reshape(world, world + 1).tolist()                         # the one read-back
    mine = cuts[rank]
    recv_counts = [cuts[s][rank + 1] - cuts[s][rank] for s in range(world)]
    total = sum(recv_counts)
    recv_idx = torch.empty(total, dtype=torch.int64, device=dev)
    recv_val = torch.empty(total, dtype=vals.dtype, device=dev)
    offs = [0]
    for c in recv_counts:
        offs.append(offs[-1] + c)
    # my own slice never leaves the GPU
    if recv_counts[rank]:
        recv_idx[offs[rank]:offs[rank + 1]].copy_(keys[mine[rank]:mine[rank + 1]])
        recv_val[offs[rank]:offs[rank + 1]].copy_(vals[mine[rank]:mine[rank + 1]])
    p2p = []
    for peer in range(world):
        if peer == rank:
            continue
        g_peer = peer if group is None else dist.get_global_rank(group, peer)
        if mine[peer + 1] > mine[peer]:
            p2p.append(dist.P2POp(dist.isend, keys[mine[peer]:mine[peer + 1]], g_peer, group))
            p2p.append(dist.P2POp(dist.isend, vals[mine[peer]:mine[peer + 1]], g_peer, group))
        if recv_counts[peer]:
            p2p.append(dist.P2POp(dist.irecv, recv_idx[offs[peer]:offs[peer + 1]], g_peer, group))
            p2p.append(dist.P2POp(dist.irecv, recv_val[offs[peer]:offs[peer + 1]], g_peer, group))
    if p2p:
        for req in dist.batch_isend_irecv(p2p):
            req.wait()
    return recv_idx, recv_val, recv_counts


def _merge_runs(ops, keys, vals, counts, op):
    """Merge the sorted runs of a received array (counts per source, in order) into one sorted,
    duplicate-free array with the union kernel: op 'add' sums values of keys present in several runs
    (partial sums of a cut row), 'override' is for runs whose keys are disjoint (transpose).  Pairwise
    tree: log2(#runs) rounds."""
    runs, off = [], 0
    for c in counts:
        if c:
            runs.append((keys[off:off + c], vals[off:off + c]))
        off += c
    if not runs:
        return keys[:0], vals[:0]
    while len(runs) > 1:
        nxt = []
        for i in range(0, len(runs) - 1, 2):
            nxt.append(ops.binop("union", op, runs[i][0], runs[i][1], runs[i + 1][0], runs[i + 1][1]))
        if len(runs) % 2:
            nxt.append(runs[-1])
        runs = nxt
    return runs[0]


_SIDE_STREAMS = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


class PeerShards:
    """A range-sharded operand whose shards live in symmetric memory: every rank can hand slices of
    ANY rank's shard to a local kernel as ordinary device pointers (peer mappings over NVLink /
    NVSwitch).  The buffers must not be modified while peers may still be reading them: ``barrier()``
    before dropping or rewriting."""

    def __init__(self, sharded):
        import torch.distributed._symmetric_memory as symm_mem
        self.shape, self.splitters = sharded.shape, list(sharded.splitters)
        self.group, self.rank, self.world = sharded.group, sharded.rank, sharded.world
        dev = sharded.indices.device
        n = torch.tensor([sharded.nnz_local], dtype=torch.int64, device=dev)
        dist.all_reduce(n, op=dist.ReduceOp.MAX, group=self.group)
        cap = max(int(n.item()), 1)                               # symmetric: the same size on every rank
        self.n_local = sharded.nnz_local
        self.dtype = sharded.data.dtype
        self._idx = symm_mem.empty(cap, dtype=torch.int64, device=dev)
        self._val = symm_mem.empty(cap, dtype=self.dtype, device=dev)
        self._idx[:self.n_local].copy_(sharded.indices)
        self._val[:self.n_local].copy_(sharded.data)
        name = (self.group or dist.group.WORLD).group_name
        self._h_idx = symm_mem.rendezvous(self._idx, name)
        self._h_val = symm_mem.rendezvous(self._val, name)
        self.barrier()

    def barrier(self):
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)

    def local_indices(self):
        return self._idx[:self.n_local]

    def local_values(self):
        return self._val[:self.n_local]

    def repartition(self, new_splitters, ops):
        """This operand on another range partition, PULLED over NVLink: every rank cuts its own shard at the new
        splitters, one small all-gather shares the cut positions (the one host read-back), and then each rank
        copies its slices straight out of its peers' shards -- plain device copies from peer-mapped memory, one
        per (source, array), starting with its right-hand neighbour so that the ranks do not all read the same
        peer at once.  No send / receive pairing, no pack or unpack; the sources are range-ordered, so the
        concatenation in source order is sorted.  Returns a ShardedFlatSparray on ``new_splitters``."""
        new_splitters = [int(x) for x in new_splitters]
        dev = self._idx.device
        bounds = torch.tensor(new_splitters, dtype=torch.int64, device=dev)
        cut = ops.lower_bound(self.local_indices(), bounds).contiguous()                 # [world + 1]
        all_cut = torch.empty(self.world * (self.world + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_cut, cut, group=self.group)
        cuts = all_cut.cpu().reshape(self.world, self.world + 1).tolist()               # the one read-back
        counts = [cuts[s][self.rank + 1] - cuts[s][self.rank] for s in range(self.world)]
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        recv_idx = torch.empty(offs[-1], dtype=torch.int64, device=dev)
        recv_val = torch.empty(offs[-1], dtype=self.dtype, device=dev)
        cur = torch.cuda.current_stream()
        side = _side_stream(dev)
        side.wait_stream(cur)
        for k in range(self.world):
            s = (self.rank + 1 + k) % self.world
            if counts[s]:
                src_idx, src_val = self.slice_of(s, cuts[s][self.rank], cuts[s][self.rank + 1])
                # my own slice is an HBM-to-HBM copy: it runs beside the NVLink pulls, on a second stream
                with torch.cuda.stream(side if s == self.rank else cur):
                    recv_idx[offs[s]:offs[s + 1]].copy_(src_idx)
                    recv_val[offs[s]:offs[s + 1]].copy_(src_val)
        recv_idx.record_stream(side)
        recv_val.record_stream(side)
        cur.wait_stream(side)
        return ShardedFlatSparray(recv_idx, recv_val, self.shape, new_splitters, self.group, ops)

    def slice_of(self, src, c0, c1):
        """(indices, values) of elements [c0, c1) of rank ``src``'s shard, as tensors over peer memory."""
        n = int(c1 - c0)
        if src == self.rank or n == 0:
            return self._idx[c0:c0 + n], self._val[c0:c0 + n]
        return (self._h_idx.get_buffer(src, (n,), torch.int64, storage_offset=int(c0)),
                self._h_val.get_buffer(src, (n,), self.dtype, storage_offset=int(c0)))


def scatter_from_full(indices, data, shape, splitters=None, group=None, ops=None, device=None):
    """Build this rank's shard from a full canonical array that every rank holds on the host
    (test / ingest helper)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if splitters is None:
        splitters = even_splitters(shape, world)
    lo = int(np.searchsorted(indices, splitters[rank], side="left"))
    hi = int(np.searchsorted(indices, splitters[rank + 1], side="left"))
    dev = device if device is not None else (ops.device() if ops is not None else CudaOps().device())
    return ShardedFlatSparray(torch.from_numpy(np.ascontiguousarray(indices[lo:hi])).to(dev),
                              torch.from_numpy(np.ascontiguousarray(data[lo:hi])).to(dev), shape, splitters, group, ops)
