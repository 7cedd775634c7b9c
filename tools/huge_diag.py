"""Stage times of the single-huge-image mode (run under torchrun): CUDA events per stage on every rank, max over ranks."""
import importlib, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "lk-s-2022-estimacija-pokreta_b200"
ops, params, synth, huge, lib = (importlib.import_module(f"{P}.{m}") for m in ("ops", "params", "synth", "huge", "_lib"))
H, W = (int(v) for v in (sys.argv[1:3] if len(sys.argv) > 2 else (2160, 3840)))
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0
D = dist if world > 1 else None
p = params.for_k(150, H=H, W=W, knn_mode=1)
a, b, _, _ = synth.make_pair(H, W, 0)
g0, g1 = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
T = {}
def timed(name, fn):
    torch.cuda.synchronize()
    if D: D.barrier()
    t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if D: D.all_reduce(t, op=dist.ReduceOp.MAX)
    T[name] = T.get(name, 0.0) + float(t.item())
    return out
for rep in range(2):
    T.clear()
    d0 = timed("daisy", lambda: ops.daisy(g0)); d1 = timed("daisy", lambda: ops.daisy(g1))
    bands = huge.band_plan(p, world); bd = bands[rank]
    sub = timed("search (own band)", lambda: ops.knn_proposals(d0[:, bd.sx0:bd.sx1].contiguous(), d1[:, bd.tx0:bd.tx1].contiguous(), huge.sub_params(p, bd))[:2])
    pv, lc, npr, lab = timed("merge (broadcast + copies + argmin)", lambda: huge.merge_bands(p, bands, rank, sub[0], sub[1], d0.device, D))
    timed("random proposals", lambda: ops.random_proposals(d0, d1, p, pv, lc, npr, lab, seed=1))
    del sub
    ws = ops.bcd_workspace(pv, world)
    kw = dict(mode=lib.BCD_INT32_F32COST, lamda=p.lamda, tpsi=p.tpsi, cost_shift=p.cost_shift)
    timed("bcd prepare (own chains)", lambda: ops.bcd_prepare(pv, lc, npr, ws, rank, world, **kw))
    for sweep in range(4):
        for ph in range(4):
            before = lab.clone()
            timed("bcd phases (own chains)", lambda: ops.bcd_phase(pv, lc, npr, lab, ws, ph, rank, world, **kw))
            def ex():
                if D:
                    delta = lab - before; D.all_reduce(delta); lab.copy_(before + delta)
            timed("label exchange", ex)
if rank == 0:
    print(f"{W}x{H} on {world} rank(s), one direction:", {k: round(v, 1) for k, v in T.items()}, "sum", round(sum(T.values()), 1))
if D: dist.destroy_process_group()
