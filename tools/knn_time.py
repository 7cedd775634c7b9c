"""CUDA-event time of the proposal search (flowb200_knn_proposals) on the bench workload, best of 5."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "lk-s-2022-estimacija-pokreta_b200"
ops, params, synth = (importlib.import_module(f"{P}.{m}") for m in ("ops", "params", "synth"))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 300
H, W = 436, 1024
p = params.for_k(K, H=H, W=W, knn_mode=1)
img1, img2, _, _ = synth.make_pair(H, W, 0)
d1, d2 = ops.daisy(torch.from_numpy(img1).cuda()), ops.daisy(torch.from_numpy(img2).cuda())
best = 1e9
for rep in range(6):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.knn_proposals(d1, d2, p)
    b.record(); b.synchronize()
    if rep: best = min(best, a.elapsed_time(b))
print(f"K={K} knn_proposals best {best:.2f} ms")
