"""BCD stage on the bench workload, piece by piece: CUDA-event time of the K-set preparation (sort + build) and of
every phase launch (32-bit chain kernel + the 64-bit fix-up launch behind it), and a full-size regression check of the
labels against the float64 implementation (bcd.cu, itself pinned against the C oracle at this size by
tests/test_gpu_baseline_configs.py).   python tools/bcd_phase_time.py [K] [sweeps] [H] [W]"""
import importlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "lk-s-2022-estimacija-pokreta_b200"
ops, params, synth, lib = (importlib.import_module(f"{P}.{m}") for m in ("ops", "params", "synth", "_lib"))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 300
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
H = int(sys.argv[3]) if len(sys.argv) > 3 else 436
W = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
p = params.for_k(K, H=H, W=W, knn_mode=1)
img1, img2, _, _ = synth.make_pair(H, W, 0)
d1, d2 = ops.daisy(torch.from_numpy(img1).cuda()), ops.daisy(torch.from_numpy(img2).cuda())
pv, lc, npr, lab = ops.knn_proposals(d1, d2, p)
ops.random_proposals(d1, d2, p, pv, lc, npr, lab, seed=1)
ws = ops.bcd_workspace(pv)
kw = dict(mode=lib.BCD_INT32_F32COST, cost_shift=p.cost_shift)


def ev():
    return torch.cuda.Event(enable_timing=True)


res = {}
for rep in range(3):
    l2 = lab.clone()
    torch.cuda.synchronize()
    e = [ev() for _ in range(2 + 4 * sweeps)]
    e[0].record()
    ops.bcd_prepare(pv, lc, npr, ws, 0, 1, **kw)
    e[1].record()
    for w in range(sweeps):
        for ph in range(4):
            ops.bcd_phase(pv, lc, npr, l2, ws, ph, 0, 1, **kw)
            e[2 + 4 * w + ph].record()
    torch.cuda.synchronize()
    t = [e[i].elapsed_time(e[i + 1]) for i in range(len(e) - 1)]
    ph = np.array(t[1:]).reshape(sweeps, 4)
    res = {"K": K, "H": H, "W": W, "sweeps": sweeps, "nprop_mean": float(npr.float().mean()),
           "prepare_ms": round(t[0], 3), "col_phase_ms": round(float(ph[:, [0, 2]].mean()), 3),
           "row_phase_ms": round(float(ph[:, [1, 3]].mean()), 3), "chains_ms": round(float(ph.sum()), 3),
           "total_ms": round(float(sum(t)), 3), "per_phase_ms": [[round(float(v), 3) for v in r] for r in ph]}
print(json.dumps(res))
whole = lab.clone()
ops.bcd(pv, lc, npr, whole, sweeps, **kw)
assert torch.equal(whole, l2), "flowb200_bcd differs from prepare + phases"
m = ops.quantise_costs(lc, p.lamda, p.cost_shift)
lq = torch.where(lc == 1000.0, torch.full_like(lc, 1000.0, dtype=torch.float64), m.double() * (20.0 / (1 << p.cost_shift)))
os.environ["FLOWB200_BCD_LEGACY"] = "1"      # bcd.cu: K-sets re-evaluated at every step (independent implementation)
ref = lab.clone()
ops.bcd(pv, lq, npr, ref, sweeps, mode=lib.BCD_FP64_F64COST)
del os.environ["FLOWB200_BCD_LEGACY"]
ok = torch.equal(ref, l2)
print("labels equal to the float64 implementation (bcd.cu):", ok, "| moved", float((l2 != lab).float().mean()))
assert ok
# the float64 programme on the K-set records (what bcd.py runs on reference-written files)
f64 = lab.clone()
torch.cuda.synchronize()
e0, e1 = ev(), ev()
e0.record()
ops.bcd(pv, lc, npr, f64, sweeps, mode=lib.BCD_FP64_F32COST)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"fp64_ksets_total_ms": round(e0.elapsed_time(e1), 3), "sweeps": sweeps}))
