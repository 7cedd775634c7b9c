"""Where the band merge of the single-huge-image mode spends its time (torchrun, 2+ ranks)."""
import importlib, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "lk-s-2022-estimacija-pokreta_b200"
params, huge = (importlib.import_module(f"{P}.{m}") for m in ("params", "huge"))
H, W = 2160, 3840
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
p = params.for_k(150, H=H, W=W, knn_mode=1)
bands = huge.band_plan(p, world); bd = bands[rank]
K = p.maxnprop
sp = torch.zeros((H, bd.width, K), dtype=torch.int32, device="cuda"); sc = torch.zeros((H, bd.width, K), dtype=torch.float32, device="cuda")
def tm(fn):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
for rep in range(3):
    (pvec, lcost), t_init = tm(lambda: (torch.full((H, W, K), -1, dtype=torch.int32, device="cuda"), torch.full((H, W, K), 1000.0, dtype=torch.float32, device="cuda")))
    t_alloc = t_bc = t_cp = 0.0
    for b in bands:
        # (every rank calls tm() the same number of times: it contains a barrier)
        (rp, rc), dt = tm(lambda: (sp, sc) if b.rank == rank else
                          (torch.empty((H, b.width, K), dtype=torch.int32, device="cuda"),
                           torch.empty((H, b.width, K), dtype=torch.float32, device="cuda")))
        bp, bc = rp, rc
        t_alloc += dt
        _, dt = tm(lambda: (dist.broadcast(bp, src=b.rank), dist.broadcast(bc, src=b.rank))); t_bc += dt
        plan = huge.copy_plan(p, b)
        def cp():
            for y0, y1, x0, x1, d0, s0, n in plan:
                pvec[y0:y1, x0:x1, d0:d0 + n] = bp[y0:y1, x0 - b.sx0:x1 - b.sx0, s0:s0 + n]
                lcost[y0:y1, x0:x1, d0:d0 + n] = bc[y0:y1, x0 - b.sx0:x1 - b.sx0, s0:s0 + n]
        _, dt = tm(cp); t_cp += dt
    if rank == 0: print(f"rep {rep}: init {t_init:.1f} alloc {t_alloc:.1f} broadcast {t_bc:.1f} copies {t_cp:.1f} ms; plan entries {sum(len(huge.copy_plan(p, b)) for b in bands)}")
dist.destroy_process_group()
