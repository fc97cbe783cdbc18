"""CPU tests of the multi-GPU layer (sparray_b200/sharded.py) with world_size 2 under gloo:
splitters, co-partitioned binops, the counts-allgather + all-to-all repartition, border
fix-up of axis sums, transpose exchange and SpMV all-reduce.  The local kernels are replaced
by a numpy/oracle stand-in INJECTED BY THE TEST (the shipped default is CudaOps; the product
has no CPU path)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))


class OracleOps:
    """Checker-backed local ops on CPU tensors (test infrastructure)."""

    def __init__(self):
        import flat_oracle as fo
        self.fo = fo

    def device(self):
        return torch.device("cpu")

    def lower_bound(self, sorted_idx, query):
        return torch.from_numpy(np.searchsorted(sorted_idx.numpy(), query.numpy(), side="left").astype(np.int64))

    def binop(self, kind, op, a_idx, a_val, b_idx, b_val):
        # "override" is only used on runs with disjoint keys, where any op that keeps x for (x, 0) will do
        uf = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "minimum": np.minimum, "maximum": np.maximum,
              "override": np.add}[op]
        ai, av, bi, bv = a_idx.numpy(), a_val.numpy(), b_idx.numpy(), b_val.numpy()
        if kind == "union":
            if op == "sub":
                i, v = self.fo.pairwise_union(ai, av, bi, -bv, np.add)
            else:
                i, v = self.fo.pairwise_union(ai, av, bi, bv, uf)
        else:
            i, v = self.fo.pairwise_intersect(ai, av, bi, bv, uf)
        return torch.from_numpy(i), torch.from_numpy(v)

    def reduce_by_key(self, keys, vals, divisor=1):
        k = keys.numpy() // divisor
        if len(k) == 0:
            return torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.float64)
        u, inv = np.unique(k, return_inverse=True)
        return torch.from_numpy(u), torch.from_numpy(np.bincount(inv, vals.numpy().astype(np.float64), minlength=len(u)))

    def sum_axis_window(self, idx, val, shape, axis, cell_lo, cell_hi, n_dev=None):
        assert n_dev is None
        i, v, _ = self.fo.sum_axis(idx.numpy(), val.numpy(), shape, axis)
        assert len(i) == 0 or (i[0] >= cell_lo and i[-1] < cell_hi), "output outside the shard's window"
        return torch.from_numpy(i), torch.from_numpy(v), None

    def can_chain_count(self, cells):
        return False

    def rekey_sort(self, idx, val, in_s, out_s, div, bits):
        k = idx.numpy().copy()
        if len(in_s):
            rem, acc = k.copy(), np.zeros_like(k)
            for g, (a, b) in enumerate(zip(in_s, out_s)):
                if g < len(in_s) - 1:
                    q = rem // a
                    rem -= q * a
                    acc += q * b
                else:
                    acc += rem * b
            k = acc
        order = np.argsort((k // div) & ((1 << bits) - 1), kind="stable") if bits else np.arange(len(k))
        return torch.from_numpy(k[order]), val[torch.from_numpy(order)]

    def axis_accumulate(self, idx, val, in_s, out_s, cells):
        k, _ = self.rekey_sort(idx, val, in_s, out_s, 1, 0)
        acc = np.zeros(cells)
        np.add.at(acc, k.numpy(), val.numpy().astype(np.float64))
        touched = np.zeros((cells + 63) // 64 * 64, dtype=np.uint8)
        touched[k.numpy()] = 1
        return torch.from_numpy(acc), torch.from_numpy(touched)

    def emit_touched(self, acc, touched, lo, hi):
        assert lo % 64 == 0 or lo == hi
        sel = np.nonzero(touched.numpy()[lo:hi])[0] + lo
        return torch.from_numpy(sel.astype(np.int64)), acc[torch.from_numpy(sel)]

    def spmv(self, idx, val, ncols, nrows, x):
        i = idx.numpy()
        y = np.bincount(i // ncols, val.numpy().astype(np.float64) * x.numpy()[i % ncols].astype(np.float64), minlength=nrows)
        return torch.from_numpy(y)


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import flat_oracle as fo
        from sparray_b200 import sharded as sh
        ops = OracleOps()
        rng = np.random.default_rng(0)          # same stream on every rank: identical "full" arrays
        shape = (12, 10, 8)
        size = int(np.prod(shape))
        ai = np.unique(rng.integers(0, size, 400)); av = rng.standard_normal(len(ai))
        bi = np.unique(rng.integers(0, size, 350)); bv = rng.standard_normal(len(bi))
        A = sh.scatter_from_full(ai, av, shape, ops=ops)
        # B arrives with DIFFERENT (sampled, unaligned) splitters -> forces the all-to-all
        odd = [0, 333, size]
        B = sh.scatter_from_full(bi, bv, shape, splitters=odd, ops=ops)
        assert A.nnz() == len(ai) and B.nnz() == len(bi)

        # co-partitioned and repartitioned binops
        for f, uf in ((lambda a, b: a + b, np.add), (lambda a, b: a - b, None), (lambda a, b: a.maximum(b), np.maximum)):
            gi, gv = f(A, B).gather()
            if uf is None:
                wi, wv = fo.pairwise_union(ai, av, bi, -bv, np.add)
            else:
                wi, wv = fo.pairwise_union(ai, av, bi, bv, uf)
            assert np.array_equal(gi, wi) and np.array_equal(gv, wv)
        gi, gv = (A * B).gather()
        wi, wv = fo.pairwise_intersect(ai, av, bi, bv, np.multiply)
        assert np.array_equal(gi, wi) and np.array_equal(gv, wv)
        R = B.repartition(A.splitters)
        assert R.splitters == A.splitters
        lo, hi = A.splitters[rank], A.splitters[rank + 1]
        ri = R.indices.numpy()
        assert np.all((ri >= lo) & (ri < hi)) and np.all(np.diff(ri) > 0)

        # sampled splitters balance the nnz
        sp = sh.sampled_splitters(A.indices, shape, ops=ops)
        assert sp[0] == 0 and sp[-1] == size and all(sp[i] <= sp[i + 1] for i in range(world))
        A2 = A.repartition(sp)
        n_loc = A2.nnz_local
        assert abs(n_loc - len(ai) / world) < 0.2 * len(ai)

        # axis sums: last axis (border fix-up, splitters inside a row), middle axis, first axis;
        # other axes through the dense accumulators + all-reduce, then again through the
        # exchange of partial sums that large outputs take
        for dense_limit in (sh.DENSE_ALLREDUCE_MAX_CELLS, 0):
            sh.DENSE_ALLREDUCE_MAX_CELLS = dense_limit
            for arr, full in ((A, (ai, av)), (A2, (ai, av)), (B, (bi, bv))):
                for axis in (2, 1, 0):
                    S = arr.sum(axis)
                    gi, gv = S.gather()
                    wi, wv, wshape = fo.sum_axis(full[0], full[1], shape, axis)
                    assert S.shape == tuple(wshape)
                    assert np.array_equal(gi, wi), (axis, gi[:10], wi[:10])
                    np.testing.assert_allclose(gv, wv, rtol=1e-12, atol=1e-12)
                    # the result obeys the splitters it advertises (round-1 review): every local key lies in
                    # this rank's range, so a co-partitioned binop on two such results is purely local
                    sk = S.indices.numpy()
                    assert np.all((sk >= S.splitters[rank]) & (sk < S.splitters[rank + 1])), (axis, S.splitters)
                    gi2, gv2 = (S + S).gather()
                    assert np.array_equal(gi2, wi)
                    np.testing.assert_allclose(gv2, 2 * wv, rtol=1e-12, atol=1e-12)
        # splitters snapped to whole blocks of the reduced axis: the sum needs no exchange at all
        for axis in (2, 1):
            block = int(np.prod(shape[axis:]))
            spa = sh.sampled_splitters(A.indices, shape, ops=ops, align=block)
            assert all(x % block == 0 for x in spa[1:-1])
            A3 = A.repartition(spa)
            S = A3.sum(axis)
            wi, wv, _ = fo.sum_axis(ai, av, shape, axis)
            gi, gv = S.gather()
            assert np.array_equal(gi, wi)
            np.testing.assert_allclose(gv, wv, rtol=1e-12, atol=1e-12)

        # transpose: re-key, exchange by new-key range, local sort
        for axes in ((2, 1, 0), (1, 0, 2), (0, 2, 1)):
            T = A2.transpose(*axes)
            gi, gv = T.gather()
            wi, wv, wshape = fo.transpose(ai, av, shape, axes)
            assert T.shape == tuple(wshape) and np.array_equal(gi, wi) and np.array_equal(gv, wv)

        # SpMV with a replicated vector
        x = rng.standard_normal(shape[-1])
        y = A2.dot(torch.from_numpy(x)).numpy()
        np.testing.assert_allclose(y, fo.dot_dense_vector(ai, av, shape, x), rtol=1e-12, atol=1e-12)

        # a shard that is empty on one rank
        ci = np.array([5, 7, 11]); cv = np.array([1.0, 2.0, 3.0])
        Cs = sh.scatter_from_full(ci, cv, shape, ops=ops)
        gi, gv = (Cs + A).gather()
        wi, wv = fo.pairwise_union(ci, cv, ai, av, np.add)
        assert np.array_equal(gi, wi) and np.array_equal(gv, wv)
        gi, gv = Cs.sum(2).gather()
        wi, wv, _ = fo.sum_axis(ci, cv, shape, 2)
        assert np.array_equal(gi, wi) and np.allclose(gv, wv)
        results[rank] = "ok"
    except Exception as exc:  # pragma: no cover
        import traceback
        results[rank] = "FAILED: %s\n%s" % (exc, traceback.format_exc())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    for r in range(world):
        assert results.get(r) == "ok", results.get(r)


def test_even_splitters_align_to_rows():
    from sparray_b200 import sharded as sh
    assert sh.even_splitters((8, 5, 4), 2) == [0, 80, 160]
    assert sh.even_splitters((3, 10), 2) == [0, 10, 30]
    s = sh.even_splitters((1, 100), 4)
    assert s == [0, 25, 50, 75, 100]
    assert sh.even_splitters((256, 256, 256, 64), 8)[1] == 32 * 256 * 256 * 64
