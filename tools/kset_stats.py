"""How many K-set records were stored vs left to the dense path, and their sizes (reads the descriptor table out of the
BCD workspace; mirrors kset_layout in csrc/bcd_ksets.cu)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "lk-s-2022-estimacija-pokreta_b200"
ops, params, synth, lib = (importlib.import_module(f"{P}.{m}") for m in ("ops", "params", "synth", "_lib"))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 300
H, W = 436, 1024
p = params.for_k(K, H=H, W=W, knn_mode=1)
img1, img2, _, _ = synth.make_pair(H, W, 0)
d1, d2 = ops.daisy(torch.from_numpy(img1).cuda()), ops.daisy(torch.from_numpy(img2).cuda())
pv, lc, npr, lab = ops.knn_proposals(d1, d2, p)
ops.random_proposals(d1, d2, p, pv, lc, npr, lab, seed=1)
ws = ops.bcd_workspace(pv)
ops.bcd_prepare(pv, lc, npr, ws, 0, 1, mode=lib.BCD_INT32_F32COST)
torch.cuda.synchronize()
al = lambda v: (v + 255) // 256 * 256
n = H * W
Kpad, Kst = (K + 31) // 32 * 32, (K + 7) // 8 * 8
col, row = (W + 1) // 2 * H, (H + 1) // 2 * W
off = al(max(col, row) * Kpad * 2) + al(n * Kst * 2) + al(n * Kst * 4)
desc = ws[off:off + 2 * n * 8].view(torch.int64).cpu().numpy()
size = (desc >> 40) << 4
stored = size > 0
print(f"K={K}: records {2 * n}, stored {stored.mean() * 100:.3f} %, dense {(~stored).sum()} "
      f"({(~stored).mean() * 100:.4f} %), mean size {size[stored].mean():.0f} B, p99 {np.percentile(size[stored], 99):.0f} B, "
      f"max {size.max()} B, arena used {size.sum() / 1e9:.2f} GB of {(ws.numel() - off - al(2 * n * 8) - 256) / 1e9:.2f} GB")
