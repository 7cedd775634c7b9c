"""Diagnostics of the BCD stage on the bench workload: CUDA-event time per sweep."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "lk-s-2022-estimacija-pokreta_b200"
ops, params, synth, lib = (importlib.import_module(f"{P}.{m}") for m in ("ops", "params", "synth", "_lib"))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 300
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
H, W = 436, 1024
p = params.for_k(K, H=H, W=W, knn_mode=1)
img1, img2, _, _ = synth.make_pair(H, W, 0)
d1, d2 = ops.daisy(torch.from_numpy(img1).cuda()), ops.daisy(torch.from_numpy(img2).cuda())
pv, lc, npr, lab = ops.knn_proposals(d1, d2, p)
ops.random_proposals(d1, d2, p, pv, lc, npr, lab, seed=1)
m = ops.quantise_costs(lc, p.lamda, p.cost_shift)
print("nprop mean", float(npr.float().mean()), "max", int(npr.max()))
for rep in range(2):
    l2 = lab.clone()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.bcd(pv, m, npr, l2, sweeps, mode=lib.BCD_INT32, cost_shift=p.cost_shift)
    b.record(); b.synchronize()
print(f"K={K} bcd {sweeps} sweep(s): {a.elapsed_time(b):.2f} ms; labels changed {(l2 != lab).float().mean().item():.3f}")
