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
  * axis sums: local segmented reduce; only output keys on a shard border can collide, so
    the borders are fixed by an allgather of one (key, value) pair per rank -- or, when the
    reduced axis is not the last one, an all-to-all of the partial sums by output-key range
    followed by a local merge-reduce;
  * transpose: re-key locally, all-to-all by new-key splitters, local sort;
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


def sampled_splitters(local_idx, shape, group=None, ops=None, samples_per_rank=256):
    """Splitters that balance nnz: every rank contributes evenly spaced samples of its sorted
    indices, the samples are all-gathered (a few KB), and the r/world quantiles become the
    boundaries."""
    world = dist.get_world_size(group)
    n = local_idx.numel()
    take = min(samples_per_rank, n)
    if take:
        pos = torch.linspace(0, n - 1, take, device=local_idx.device).long()
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
    cuts = [0] + [int(allv[len(allv) * r // world]) for r in range(1, world)] + [size]
    for r in range(1, world + 1):          # keep them monotone
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts


class ShardedFlatSparray:
    """This rank's shard of a range-partitioned FlatSparray."""

    def __init__(self, indices, data, shape, splitters, group=None, ops=None):
        self.indices, self.data = indices, data
        self.shape = tuple(int(s) for s in shape)
        self.splitters = [int(s) for s in splitters]
        self.group = group
        self.ops = ops if ops is not None else CudaOps()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        assert len(self.splitters) == self.world + 1

    @property
    def nnz_local(self):
        return int(self.indices.numel())

    def nnz(self):
        t = torch.tensor([self.nnz_local], dtype=torch.int64, device=self.indices.device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def _like(self, indices, data, shape=None, splitters=None):
        return ShardedFlatSparray(indices, data, shape or self.shape, splitters or self.splitters,
                                  self.group, self.ops)

    # ---- co-partitioned / repartitioned binops -------------------------------------------------------
    def _binop(self, other, kind, op):
        assert self.shape == other.shape
        if other.splitters != self.splitters:
            other = other.repartition(self.splitters)       # the only communication of a binop
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

    # ---- repartition: counts allgather + one variable-size all-to-all --------------------------------
    def repartition(self, new_splitters):
        new_splitters = [int(s) for s in new_splitters]
        dev = self.indices.device
        # destination d receives my elements in [new[d], new[d+1]): a contiguous slice of my sorted shard
        bounds = torch.tensor(new_splitters, dtype=torch.int64, device=dev)
        cut = self.ops.lower_bound(self.indices, bounds).cpu().tolist()
        send_counts = [cut[d + 1] - cut[d] for d in range(self.world)]
        counts = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        all_counts = [torch.empty_like(counts) for _ in range(self.world)]
        dist.all_gather(all_counts, counts, group=self.group)
        recv_counts = torch.stack(all_counts)[:, self.rank].cpu().tolist()      # one read-back, not one per source
        lo = cut[0]
        send_idx = self.indices[lo:cut[-1]].contiguous()
        send_val = self.data[lo:cut[-1]].contiguous()
        recv_idx = torch.empty(sum(recv_counts), dtype=torch.int64, device=dev)
        recv_val = torch.empty(sum(recv_counts), dtype=self.data.dtype, device=dev)
        dist.all_to_all_single(recv_idx, send_idx, recv_counts, send_counts, group=self.group)
        dist.all_to_all_single(recv_val, send_val, recv_counts, send_counts, group=self.group)
        # sources are range-ordered, so the concatenation in source order is globally sorted
        return self._like(recv_idx, recv_val, splitters=new_splitters)

    # ---- reductions -------------------------------------------------------------------------------------
    def sum(self, axis):
        """sum over one axis -> ShardedFlatSparray with float64 data (flat.py:607-626)."""
        from . import _plan
        axis = int(axis) % len(self.shape)
        new_shape = self.shape[:axis] + self.shape[axis + 1:]
        plan = _plan.axis_sum_plan(self.shape, axis)
        dev = self.indices.device
        if plan[0] == "last":
            div = plan[1]
            keys, sums = self.ops.reduce_by_key(self.indices, self.data, div)
            new_split = [s // div for s in self.splitters[:-1]] + [_size(new_shape)]
            return self._fix_borders(keys, sums, new_shape, new_split)
        _, in_s, out_s, bits = plan
        cells = _size(new_shape)
        if cells <= DENSE_ALLREDUCE_MAX_CELLS:
            # reduced array small enough to hold densely on every GPU: local float64
            # accumulators, one all-reduce of the sums and one of the "touched" flags, then every rank
            # emits its own range of output cells (ranges aligned to 64 cells)
            acc, touched = self.ops.axis_accumulate(self.indices, self.data, in_s, out_s, cells)
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(touched, op=dist.ReduceOp.MAX, group=self.group)
            per = -(-cells // self.world)
            per = -(-per // 64) * 64
            out_split = [min(r * per, cells) for r in range(self.world)] + [cells]
            keys, sums = self.ops.emit_touched(acc, touched, out_split[self.rank], out_split[self.rank + 1])
            return ShardedFlatSparray(keys, sums, new_shape, out_split, self.group, self.ops)
        # middle axis, large output: local re-key + sort + reduce, then exchange partial sums by output-key range
        keys, vals = self.ops.rekey_sort(self.indices, self.data, in_s, out_s, 1, bits)
        keys, sums = self.ops.reduce_by_key(keys, vals, 1)
        out_split = even_splitters(new_shape, self.world)
        part = ShardedFlatSparray(keys, sums, new_shape, [0] * self.world + [_size(new_shape)], self.group, self.ops)
        part.splitters = None                                 # not range-partitioned yet
        got = _exchange_unsorted(part, out_split)
        # every source contributed a sorted run: sort by key again (bits of the output key) and reduce
        kbits = max(_size(new_shape) - 1, 1).bit_length()
        k2, v2 = self.ops.rekey_sort(got[0], got[1], [], [], 1, kbits)
        k3, s3 = self.ops.reduce_by_key(k2, v2, 1)
        return ShardedFlatSparray(k3, s3, new_shape, out_split, self.group, self.ops)

    def _fix_borders(self, keys, sums, new_shape, new_split):
        """After a last-axis reduction only the FIRST output key of a rank can also be the last
        key of an earlier rank (a shard boundary inside a row).  Everybody all-gathers
        (first key, last key, count, first sum); the earliest rank of each chain keeps the key
        and absorbs the partial sums, the others drop their first element."""
        dev = keys.device
        n = keys.numel()
        meta = torch.tensor([int(keys[0]) if n else -1, int(keys[-1]) if n else -1, n], dtype=torch.int64, device=dev)
        info = torch.zeros(1, dtype=torch.float64, device=dev)
        if n:
            info[0] = sums[0]
        metas = [torch.empty_like(meta) for _ in range(self.world)]
        infos = [torch.empty_like(info) for _ in range(self.world)]
        dist.all_gather(metas, meta, group=self.group)
        dist.all_gather(infos, info, group=self.group)
        firsts = [int(m[0]) for m in metas]
        lasts = [int(m[1]) for m in metas]
        counts = [int(m[2]) for m in metas]
        dup = [False] * self.world
        owner = list(range(self.world))
        prev = None
        for r in range(self.world):
            if counts[r] == 0:
                continue
            if prev is not None and lasts[prev] == firsts[r]:
                dup[r] = True
                owner[r] = owner[prev] if (counts[prev] == 1 and dup[prev]) else prev
            prev = r
        sums = sums.clone()
        if dup[self.rank]:
            keys, sums = keys[1:], sums[1:]
        for r in range(self.rank + 1, self.world):
            if dup[r] and owner[r] == self.rank:
                sums[-1] += infos[r][0].to(sums.device)
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
        from . import _plan
        ndim = len(self.shape)
        axes = tuple(range(ndim - 1, -1, -1)) if not axes else tuple(int(a) for a in axes)
        new_shape = tuple(self.shape[a] for a in axes)
        in_s, out_s, _, _ = _plan.transpose_plan(self.shape, axes)
        keys, vals = self.ops.rekey_sort(self.indices, self.data, in_s, out_s, 1, 0)     # re-key only
        out_split = even_splitters(new_shape, self.world)
        part = ShardedFlatSparray(keys, vals, new_shape, [0] * (self.world + 1), self.group, self.ops)
        got = _exchange_unsorted(part, out_split)
        kbits = max(_size(new_shape) - 1, 1).bit_length()
        k2, v2 = self.ops.rekey_sort(got[0], got[1], [], [], 1, kbits)
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

    def slice_of(self, src, c0, c1):
        """(indices, values) of elements [c0, c1) of rank ``src``'s shard, as tensors over peer memory."""
        n = int(c1 - c0)
        if src == self.rank or n == 0:
            return self._idx[c0:c0 + n], self._val[c0:c0 + n]
        return (self._h_idx.get_buffer(src, (n,), torch.int64, storage_offset=int(c0)),
                self._h_val.get_buffer(src, (n,), self.dtype, storage_offset=int(c0)))


def _exchange_unsorted(part, out_split):
    """All-to-all of (key, value) pairs whose keys are NOT range-partitioned by source (after a
    re-key): the pairs are bucketed by destination with one stable sort pass on the
    destination id, exchanged, and returned unsorted (the caller sorts locally)."""
    ops, group = part.ops, part.group
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    keys, vals = part.indices, part.data
    dev = keys.device
    bounds = torch.tensor(out_split[1:-1], dtype=torch.int64, device=dev)
    # destination of every pair = number of boundaries <= key: keys are sorted locally after the
    # caller's reduce / (for transpose) not sorted -- so sort them first by the full key
    kbits = max(int(out_split[-1]) - 1, 1).bit_length()
    keys, vals = ops.rekey_sort(keys, vals, [], [], 1, kbits)
    full = torch.tensor(out_split, dtype=torch.int64, device=dev)
    cut = ops.lower_bound(keys, full).cpu().tolist()
    send_counts = [cut[d + 1] - cut[d] for d in range(world)]
    counts = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    recv_counts = [int(all_counts[s][rank].item()) for s in range(world)]
    recv_idx = torch.empty(sum(recv_counts), dtype=torch.int64, device=dev)
    recv_val = torch.empty(sum(recv_counts), dtype=vals.dtype, device=dev)
    dist.all_to_all_single(recv_idx, keys.contiguous(), recv_counts, send_counts, group=group)
    dist.all_to_all_single(recv_val, vals.contiguous(), recv_counts, send_counts, group=group)
    return recv_idx, recv_val


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
