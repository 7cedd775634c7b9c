"""Diagnostics of the kNN stage on the bench workload: per-kernel CUDA-event times and tcgen05 counters."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "lk-s-2022-estimacija-pokreta_b200"
ops, params, synth = (importlib.import_module(f"{P}.{m}") for m in ("ops", "params", "synth"))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 300
H, W = 436, 1024
p = params.for_k(K, H=H, W=W, knn_mode=1)
img1, img2, _, _ = synth.make_pair(H, W, 0)
d1, d2 = ops.daisy(torch.from_numpy(img1).cuda()), ops.daisy(torch.from_numpy(img2).cuda())
for rep in range(2):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if os.environ.get("KNN_DIAG_STATS", "1") == "1":
        pv, lc, npr, lab, idx, stats = ops.knn_proposals(d1, d2, p, want_idx=True)
    else:
        pv, lc, npr, lab = ops.knn_proposals(d1, d2, p)
        stats = torch.zeros(8, dtype=torch.int32)
    b.record(); b.synchronize()
    st = stats.cpu().numpy().astype(np.int64)
ntask = int((npr.cpu().numpy() // p.k_cell).sum())
print(f"K={K} k_cell={p.k_cell} knn {a.elapsed_time(b):.2f} ms; tasks {ntask}; fallback {st[0]} collect-ovf {st[1]} cand-ovf {st[2]}; "
      f"mean cand {(st[4] & 0xffffffff) / max(1, ntask):.2f}")
