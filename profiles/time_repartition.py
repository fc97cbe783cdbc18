"""Phase timing of ShardedFlatSparray.repartition under torchrun (NCCL): where the ~1 ms goes.
Usage: torchrun --nproc-per-node 2 profiles/time_repartition.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from sparray_b200 import sharded as sh

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ai, av, bi, bv = bench.gen_operands(rank)
size = int(np.prod(bench.SHAPE))
gshape = (bench.SHAPE[0] * world,) + tuple(bench.SHAPE[1:])
even = [r * size for r in range(world + 1)]
shifted = [0] + [r * size + size // 2 for r in range(world - 1)] + [world * size]
B = sh.ShardedFlatSparray(torch.from_numpy(bi).cuda() + rank * size, torch.from_numpy(bv).cuda(), gshape, even)
Bs = B.repartition(shifted)
for _ in range(3):
    Bs.repartition(even)
torch.cuda.synchronize(); dist.barrier()

def phase_times():
    self = Bs
    t = [time.perf_counter()]
    dev = self.indices.device
    bounds = torch.tensor(even, dtype=torch.int64, device=dev)
    cut = self.ops.lower_bound(self.indices, bounds).cpu().tolist(); t.append(time.perf_counter())
    send_counts = [cut[d + 1] - cut[d] for d in range(world)]
    counts = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    recv_counts = torch.stack(all_counts)[:, rank].cpu().tolist(); t.append(time.perf_counter())
    send_idx = self.indices[cut[0]:cut[-1]]; send_val = self.data[cut[0]:cut[-1]]
    recv_idx = torch.empty(sum(recv_counts), dtype=torch.int64, device=dev)
    recv_val = torch.empty(sum(recv_counts), dtype=self.data.dtype, device=dev)
    if os.environ.get("VAL_FIRST"):
        dist.all_to_all_single(recv_val, send_val, recv_counts, send_counts)
        torch.cuda.synchronize(); t.append(time.perf_counter())
        dist.all_to_all_single(recv_idx, send_idx, recv_counts, send_counts)
        torch.cuda.synchronize(); t.append(time.perf_counter())
    else:
        dist.all_to_all_single(recv_idx, send_idx, recv_counts, send_counts)
        torch.cuda.synchronize(); t.append(time.perf_counter())
        dist.all_to_all_single(recv_val, send_val, recv_counts, send_counts)
        torch.cuda.synchronize(); t.append(time.perf_counter())
    return [(b - a) * 1e3 for a, b in zip(t[:-1], t[1:])]

def list_variant():
    self = Bs
    dev = self.indices.device
    bounds = torch.tensor(even, dtype=torch.int64, device=dev)
    cut = self.ops.lower_bound(self.indices, bounds).cpu().tolist()
    send_counts = [cut[d + 1] - cut[d] for d in range(world)]
    counts = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    recv_counts = torch.stack(all_counts)[:, rank].cpu().tolist()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    si = [self.indices[cut[d]:cut[d + 1]] for d in range(world)]
    sv = [self.data[cut[d]:cut[d + 1]] for d in range(world)]
    ri = [torch.empty(c, dtype=torch.int64, device=dev) for c in recv_counts]
    rv = [torch.empty(c, dtype=self.data.dtype, device=dev) for c in recv_counts]
    dist.all_to_all(ri, si)
    dist.all_to_all(rv, sv)
    out_i, out_v = torch.cat(ri), torch.cat(rv)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3

lv = 0.0
for _ in range(3):
    list_variant()
for _ in range(10):
    dist.barrier()
    lv += list_variant()
if rank == 0:
    print("list all_to_all (idx + val + cat): %.3f ms" % (lv / 10))
acc = np.zeros(4)
for _ in range(10):
    dist.barrier()
    acc += np.array(phase_times())
if rank == 0:
    print("ms: lower_bound+readback %.3f | counts all_gather+readback %.3f | all_to_all idx %.3f | all_to_all val %.3f" % tuple(acc / 10))
dist.destroy_process_group()
