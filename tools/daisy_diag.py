"""DAISY on the bench image: CUDA-event time per image and algorithmic GB/s (275 B/pixel)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "lk-s-2022-estimacija-pokreta_b200"
ops, synth = (importlib.import_module(f"{P}.{m}") for m in ("ops", "synth"))
H, W = 436, 1024
g = torch.from_numpy(synth.texture(H, W, 3)).cuda()
for _ in range(3): ops.daisy(g)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): ops.daisy(g)
b.record(); b.synchronize()
ms = a.elapsed_time(b) / 20
print(f"daisy {ms:.4f} ms/image, {H * W * 275 / ms / 1e6:.0f} GB/s algorithmic")
