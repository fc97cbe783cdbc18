"""Feasibility probe: torch symmetric memory + our kernels reading a PEER shard over NVLink.
torchrun --nproc-per-node 2 profiles/probe_symm.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from sparray_b200 import _kernels as k

n = 21_400_000
g = torch.Generator(device="cuda"); g.manual_seed(7)            # same stream on both ranks
steps = torch.randint(1, 20, (n,), device="cuda", generator=g, dtype=torch.int64)
full = torch.cumsum(steps, 0)
vals = (full % 97).to(torch.float32)
# rank r publishes ITS half of b in symmetric memory
half = n // 2
mine = slice(rank * half, (rank + 1) * half)
sym_idx = symm_mem.empty(half, dtype=torch.int64, device="cuda")
sym_val = symm_mem.empty(half, dtype=torch.float32, device="cuda")
sym_idx.copy_(full[mine]); sym_val.copy_(vals[mine])
h_idx = symm_mem.rendezvous(sym_idx, dist.group.WORLD.group_name)
h_val = symm_mem.rendezvous(sym_val, dist.group.WORLD.group_name)
h_idx.barrier()
peer = 1 - rank
p_idx = h_idx.get_buffer(peer, (half,), torch.int64)
p_val = h_val.get_buffer(peer, (half,), torch.float32)
# a = every third element of the PEER's half (so a & b_peer == a), held locally
a_idx = full[peer * half:(peer + 1) * half][::3].contiguous()
a_val = (a_idx % 89).to(torch.float32)
torch.cuda.synchronize()
oi, ov, cnt = k.setop_binop("intersect", "mul", a_idx, a_val, p_idx, p_val, lazy=True)
c = int(cnt.item())
ok = c == a_idx.numel() and bool((oi[:c] == a_idx).all()) and bool((ov[:c] == a_val * (a_idx % 97).to(torch.float32)).all())

def timed(bi, bv, reps=20):
    for _ in range(3):
        k.setop_binop("intersect", "mul", a_idx, a_val, bi, bv, lazy=True)
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); k.setop_binop("intersect", "mul", a_idx, a_val, bi, bv, lazy=True); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
t_peer = timed(p_idx, p_val)
l_idx, l_val = p_idx.clone(), p_val.clone()              # the same slice copied to local HBM
t_local = timed(l_idx, l_val)
print("rank %d: intersection of a local a (%d) with a b (%d) read from the PEER: %s; device time %.1f us "
      "(%.0f GB/s of index keys over NVLink) vs %.1f us with b in local HBM" % (
          rank, a_idx.numel(), half, "OK" if ok else "WRONG", t_peer * 1e3, half * 8 / (t_peer * 1e-3) / 1e9, t_local * 1e3))
h_idx.barrier()
dist.destroy_process_group()
