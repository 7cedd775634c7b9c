"""Distribution of the K-set list lengths |S_l| on the bench workload (sampled pixel pairs, torch on the GPU; diagnostic)."""
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
rng = np.random.default_rng(0)
lens, maxes, sums, ns = [], [], [], []
for _ in range(3000):
    y, x = int(rng.integers(1, H)), int(rng.integers(1, W))
    for (qy, qx) in ((y - 1, x), (y, x - 1)):
        n, nq = int(npr[y, x]), int(npr[qy, qx])
        v, u = pv[y, x, :n], pv[qy, qx, :nq]
        dy = ((v << 16) >> 16)[:, None] - ((u << 16) >> 16)[None, :]
        dx = (v >> 16)[:, None] - (u >> 16)[None, :]
        c = ((dy.abs() + dx.abs()) < 8).sum(1)
        lens.append(c.cpu().numpy()); maxes.append(int(c.max())); sums.append(int(c.sum())); ns.append(n)
l = np.concatenate(lens)
print("labels", len(l), "mean len", l.mean(), "zero frac", (l == 0).mean())
print("percentiles 50/75/90/95/99/99.9:", [int(np.percentile(l, q)) for q in (50, 75, 90, 95, 99, 99.9)], "max", l.max())
print("groups of 4: mean", np.ceil(l / 4).mean(), " groups>2 frac", (l > 8).mean(), " groups>4 frac", (l > 16).mean())
m = np.array(maxes)
print("per-record max len: mean", m.mean(), "p50/p90/p99/max", [int(np.percentile(m, q)) for q in (50, 90, 99, 100)])
print("per-record entries: mean", np.mean(sums), "mean n", np.mean(ns))
# warp-level cost model: labels sorted by decreasing length, warps of 32: sum over warps of max groups; and critical path
tot, crit = [], []
for c in lens:
    s = np.sort(c)[::-1]
    g = np.ceil(s / 4)
    w = [g[i:i + 32].max() for i in range(0, len(g), 32)]
    tot.append(sum(w)); crit.append(max(w))
print("warp-groups per record: mean", np.mean(tot), " critical (longest) groups: mean", np.mean(crit), "p90", np.percentile(crit, 90))
