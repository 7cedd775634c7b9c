"""Times edge.canny_ivice on the device (flowb200_canny_edges) on the bench workload's frames and checks every result
against cv2 itself.  Writes gpurun_out/edges_time.json.  (cv2 / the oracle are used as the checker only.)"""
import importlib, json, os, sys, time
import cv2, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
P = "lk-s-2022-estimacija-pokreta_b200"
ops, synth, lib = (importlib.import_module(f"{P}.{m}") for m in ("ops", "synth", "_lib"))

out = {"fields": {}}
rng = np.random.default_rng(3)
for (H, W) in ((436, 1024), (2160, 3840)):
    img1, img2, _, _ = synth.make_pair(H, W, 0)
    noisy = np.clip(img2.astype(np.int32) + rng.integers(-40, 41, size=img2.shape), 0, 255).astype(np.uint8)
    for name, img, (lo, hi) in (("frame", img1, (100, 200)), ("noisy_frame", noisy, (100, 200)), ("noisy_low", noisy, (20, 60))):
        t0 = time.perf_counter()
        bl = cv2.GaussianBlur(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), (3, 3), 0)
        want = np.array((255 - cv2.Canny(image=bl, threshold1=lo, threshold2=hi)) / 255, dtype="float32")
        cpu_ms = (time.perf_counter() - t0) * 1e3
        g = torch.from_numpy(img).cuda()
        ws = torch.empty(int(lib.load().flowb200_edges_workspace_bytes(H, W)), dtype=torch.uint8, device="cuda")
        ts = []
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            got = ops.canny_edges(g, lo, hi, workspace=ws)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts[1:])
        out["fields"][f"{W}x{H}_{name}_{lo}_{hi}"] = {
            "gpu_ms": round(ms, 4), "cv2_cpu_ms": round(cpu_ms, 3), "edge_pixels": int((want == 0).sum()),
            "equal_to_cv2": bool(np.array_equal(got.cpu().numpy(), want)),
            "algorithmic_GBps": round(H * W * 7 / ms / 1e6, 1)}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "edges_time.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out))
