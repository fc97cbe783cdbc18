"""Multi-GPU parity check of sparray_b200/sharded.py on real GPUs over NCCL (not a pytest
file: launch with torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py

Every rank builds the same full arrays on the host (seeded), takes its shard, runs the
sharded ops through the CUDA kernels and compares the gathered result with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import flat_oracle as fo
    from sparray_b200 import sharded as sh
    rng = np.random.default_rng(0)
    shape = (64, 48, 40, 16)
    size = int(np.prod(shape))
    ai = np.unique(rng.integers(0, size, 300000)); av = rng.standard_normal(len(ai)).astype(np.float32)
    bi = np.unique(rng.integers(0, size, 280000)); bv = rng.standard_normal(len(bi)).astype(np.float32)
    A = sh.scatter_from_full(ai, av, shape)
    sp = sh.sampled_splitters(A.indices, shape)
    odd = sorted(set([0, size] + [int(x) for x in rng.integers(1, size, world - 1)]))
    while len(odd) < world + 1:
        odd.insert(1, odd[1])
    B = sh.scatter_from_full(bi, bv, shape, splitters=odd)

    def same(got, want, exact=True):
        gi, gv = got.gather()
        assert np.array_equal(gi, want[0]), "indices differ"
        if exact:
            assert np.array_equal(gv, want[1]), "values differ"
        else:
            np.testing.assert_allclose(gv, want[1], rtol=1e-5, atol=1e-5 * np.abs(want[1]).max())

    same(A + B, fo.pairwise_union(ai, av, bi, bv, np.add))                    # all-to-all repartition of B
    same(A * B, fo.pairwise_intersect(ai, av, bi, bv, np.multiply))
    same(A - B, fo.pairwise_union(ai, av, bi, -bv, np.add))
    A2 = A.repartition(sp)
    same(A2 + B, fo.pairwise_union(ai, av, bi, bv, np.add))
    # the same misaligned binops with NO exchange step: B's shards in symmetric memory, the merge
    # kernels read the partner's slices over NVLink
    Bp = B.publish()
    # the pulled repartition (peer copies over NVLink) against the NCCL one
    for target in (A.splitters, sp, odd):
        got, want = Bp.repartition(target, B.ops), B.repartition(target)
        assert got.splitters == want.splitters
        assert torch.equal(got.indices, want.indices) and torch.equal(got.data, want.data), "pulled repartition differs"
    for X in (A, A2):
        same(X.binop_peer(Bp, "union", "add"), fo.pairwise_union(ai, av, bi, bv, np.add))
        same(X.binop_peer(Bp, "intersect", "mul"), fo.pairwise_intersect(ai, av, bi, bv, np.multiply))
        same(X.binop_peer(Bp, "union", "sub"), fo.pairwise_union(ai, av, bi, -bv, np.add))
        # a published operand handed to the ordinary operators: pulled repartition + local merge
        same(X + Bp, fo.pairwise_union(ai, av, bi, bv, np.add))
        same(X * Bp, fo.pairwise_intersect(ai, av, bi, bv, np.multiply))
    Bp.barrier()
    for axis in range(4):
        S = A2.sum(axis)
        same(S, fo.sum_axis(ai, av, shape, axis)[:2], exact=False)
        sk = S.indices
        assert bool(((sk >= S.splitters[rank]) & (sk < S.splitters[rank + 1])).all().item()), "sum() result breaks its splitters"
    # the bench chain on co-partitioned, block-aligned shards: no collective, no host wait until the end
    for axis in (2, 3, 1):
        block = int(np.prod(shape[axis:]))
        spa = sh.sampled_splitters(A.indices, shape, align=block)
        A3, B3 = A.repartition(spa), B.repartition(spa)
        C = A3 * B3
        S = C.sum(axis)
        ci, cv = fo.pairwise_intersect(ai, av, bi, bv, np.multiply)
        same(S, fo.sum_axis(ci, cv, shape, axis)[:2], exact=False)
        same(C, (ci, cv))
        U = A3 + B3
        ui, uv = fo.pairwise_union(ai, av, bi, bv, np.add)
        same(U.sum(axis), fo.sum_axis(ui, uv, shape, axis)[:2], exact=False)
    for axes in ((3, 2, 1, 0), (2, 0, 3, 1)):
        same(A2.transpose(*axes), fo.transpose(ai, av, shape, axes)[:2])
    x = rng.standard_normal(shape[-1]).astype(np.float32)
    y = A2.dot(torch.from_numpy(x).cuda()).cpu().numpy()
    np.testing.assert_allclose(y, fo.dot_dense_vector(ai, av, shape, x), rtol=1e-5, atol=1e-5)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("dist_gpu_check ok on %d GPUs" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
