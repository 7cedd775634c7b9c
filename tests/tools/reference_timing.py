"""Times the reference's OWN, unmodified Python (`python bcd.py`, `postprocessing.py`) on the small golden case, in the
build container (it needs /root/reference, so it cannot run on the GPU box), next to the C oracle on the same inputs.
SURVEY 8(d) asks for this data point: ms per pixel per sweep of the real reference, to put the oracle-port CPU
baseline of bench.py into proportion.  Writes profiles/r01_reference_python_timing.json."""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_case, load_npz
from oracle import cport, ref_harness as rh

assert rh.available()
z = load_case("pair_a")
H, W = int(z["meta"][0]), int(z["meta"][1])
d = tempfile.mkdtemp()
rh.write_stage1_files(d, 3, 0, z["b0_proposals"], z["b0_lcosts"], z["b0_nprop"], z["b0_labels00"], z["b0_packedksets"])
t0 = time.perf_counter()
labels, _ = rh.run_stage2(d, H, W, 3, 0, 1)
t_ref = time.perf_counter() - t0
assert np.array_equal(labels[0], z["b0_labels01"])
cport.set_num_threads(1)
t0 = time.perf_counter()
for _ in range(20):
    lab = cport.ceo_bcd(z["b0_proposals"], z["b0_lcosts"], z["b0_nprop"], z["b0_labels00"], 1)
t_c = (time.perf_counter() - t0) / 20
assert np.array_equal(lab[0], z["b0_labels01"])

pp = rh.reference_postprocessing()
zc = load_npz("consistency")
np.save(os.path.join(d, "f.npy"), zc["real_fwd"]); np.save(os.path.join(d, "b.npy"), zc["real_bwd"])
t0 = time.perf_counter()
pp.postProcessing(os.path.join(d, "f.npy"), os.path.join(d, "b.npy"), 5, os.path.join(d, "o.npy"))
t_pp = time.perf_counter() - t0
npix_pp = zc["real_fwd"].shape[0] * zc["real_fwd"].shape[1]

full = 436 * 1024
out = {
    "where": f"build container, {os.cpu_count()} logical CPUs, single-threaded Python (the reference is single-threaded)",
    "bcd": {"case": f"tests/golden/pair_a: {H}x{W}, K=150, 1 sweep, labels equal to the golden ones",
            "reference_python_s": round(t_ref, 3), "reference_ms_per_pixel_per_sweep": round(t_ref * 1e3 / (H * W), 3),
            "extrapolated_1024x436_one_sweep_hours": round(t_ref / (H * W) * full / 3600, 2),
            "c_oracle_1_thread_s": round(t_c, 5), "c_oracle_us_per_pixel_per_sweep": round(t_c * 1e6 / (H * W), 2)},
    "postProcessing": {"case": f"tests/golden/consistency 'real': {npix_pp} pixels incl. np.load / np.save",
                       "reference_python_s": round(t_pp, 4), "reference_us_per_pixel": round(t_pp * 1e6 / npix_pp, 1)},
}
with open(os.path.join(ROOT, "profiles", "r01_reference_python_timing.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1))
