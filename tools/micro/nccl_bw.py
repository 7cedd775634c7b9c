"""NCCL bandwidth on this box (torchrun): broadcast / all_gather / send-recv of a large int32 tensor."""
import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = 1 << 30   # 4 GiB of int32
x = torch.full((n,), rank, dtype=torch.int32, device="cuda")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / reps
tb = t(lambda: dist.broadcast(x, src=0))
out = torch.empty((world * (n // 4),), dtype=torch.int32, device="cuda")
tg = t(lambda: dist.all_gather_into_tensor(out, x[: n // 4]))
def sr():
    if rank == 0: dist.send(x, dst=1)
    elif rank == 1: dist.recv(x, src=0)
ts = t(sr)
if rank == 0:
    print(f"world {world}: broadcast 4 GiB {4.295 / tb:.1f} GB/s; all_gather (1 GiB per rank) {world * 1.074 / tg:.1f} GB/s out; send/recv 4 GiB {4.295 / ts:.1f} GB/s")
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
