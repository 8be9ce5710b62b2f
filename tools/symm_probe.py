"""Development aid (torchrun, >= 2 GPUs): can a kernel of rank r store straight into rank 0's HBM?
Probes torch.distributed._symmetric_memory (peer pointers over NVLink) with the library's own
search kernel writing (L, R) into the peer buffer."""
import ctypes as C
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem

n = 1 << 20
t = symm_mem.empty(n * world, dtype=torch.int32, device=dev)
t.fill_(-1)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
peer0 = hdl.get_buffer(0, (n * world,), torch.int32)
# every rank writes its slice into rank 0's buffer
src = torch.full((n,), rank, dtype=torch.int32, device=dev)
hdl.barrier()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(10):
    peer0[rank * n:(rank + 1) * n].copy_(src)
hdl.barrier()
torch.cuda.synchronize()
dt = (time.time() - t0) / 10
if rank == 0:
    ok = all(int(t[r * n]) == r and int(t[(r + 1) * n - 1]) == r for r in range(world))
    print("peer stores visible on rank 0:", ok, f"{dt*1e6:.0f} us per round", flush=True)
dist.barrier()
dist.destroy_process_group()
