"""Developer probe (torchrun, N GPUs): cost of ncclAllGather of the packed root statistics vs torch symmetric memory
(peer copies + barrier, multicast availability)."""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
chunk = 4096 * 5 * 4 + 4096 * 8  # config 2: visit counts + root values of 4,096 roots
local = torch.zeros(chunk, dtype=torch.uint8, device=dev)
allb = torch.zeros(world * chunk, dtype=torch.uint8, device=dev)

def timeit(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

t_nccl = timeit(lambda: dist.all_gather_into_tensor(allb, local))
msg = "nccl all_gather %d B x %d: %.1f us" % (chunk, world, t_nccl)
try:
    sb = symm_mem.empty(world * chunk, dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(sb, dist.group.WORLD)
    peers = [hdl.get_buffer(r, (world * chunk,), torch.uint8) for r in range(world)]
    mine = sb[rank * chunk:(rank + 1) * chunk]
    def put():
        for r in range(world):
            if r != rank:
                peers[r][rank * chunk:(rank + 1) * chunk].copy_(mine, non_blocking=True)
    t_bar = timeit(lambda: hdl.barrier(channel=0))
    t_put = timeit(put)
    def both():
        put(); hdl.barrier(channel=0)
    t_both = timeit(both)
    msg += "; symm_mem barrier %.1f us, %d peer copies %.1f us, put + barrier %.1f us; multicast support %s" % (
        t_bar, world - 1, t_put, t_both, hdl.has_multicast_support)
    # correctness
    mine.fill_(rank + 1); torch.cuda.synchronize(); dist.barrier()
    both(); torch.cuda.synchronize(); dist.barrier()
    ok = all(int(sb[r * chunk].item()) == r + 1 and int(sb[(r + 1) * chunk - 1].item()) == r + 1 for r in range(world))
    msg += "; gathered ok %s; buffer_ptrs %d" % (ok, len(hdl.buffer_ptrs))
except Exception as e:
    msg += "; symm_mem failed: %r" % (e,)
if rank == 0:
    print(msg, flush=True)
dist.destroy_process_group()
