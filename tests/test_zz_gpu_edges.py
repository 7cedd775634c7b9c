"""edge.canny_ivice (edge.py:19-35) on the GPU vs the reference's own outputs, the oracle and cv2.
Bit-exact edge maps.  (Named to run last: after the hot-path GPU tests and the segment tests.)"""
import importlib
import os
import sys

import cv2
import numpy as np
import pytest

from helpers import load_npz, pkg
from oracle import edges as oe

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ("texture", "noise", "boxes", "ramp", "row", "column")


def test_canny_ivice_module_files_golden(tmp_path):
    """import edge; canny_ivice(png, bin): raw float32 [H][W], 0.0 on an edge, as the reference's function wrote it."""
    sys.path.insert(0, ROOT)
    edge = importlib.import_module("edge")
    z = load_npz("edges")
    for case in CASES:
        png, binf = str(tmp_path / (case + ".png")), str(tmp_path / (case + ".bin"))
        assert cv2.imwrite(png, z[case + "_img"])
        assert edge.canny_ivice(png, binf) is None
        e = np.fromfile(binf, dtype=np.float32).reshape(z[case + "_img"].shape[:2])
        assert set(np.unique(e)) <= {0.0, 1.0}
        assert np.array_equal(e == 0, z[case + "_edges"] == 1), case
    with pytest.raises(NotImplementedError):
        edge.sed_ivice("a.png", "a.bin")


@pytest.mark.parametrize("shape", [(57, 83), (1, 1), (2, 3), (1, 200), (150, 1), (128, 256)])
def test_canny_edges_vs_oracle_random(shape):
    ops = pkg("ops")
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    for it, (lo, hi) in enumerate(((100, 200), (30, 90), (5, 10), (0, 0))):
        img = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
        if it % 2:
            img = cv2.GaussianBlur(img, (5, 5), 2.0) if min(shape) > 1 else img
        got = ops.canny_edges(torch.from_numpy(img).cuda(), lo, hi).cpu().numpy()
        assert np.array_equal(got, oe.canny_ivice(img, lo, hi)), (it, lo, hi)


def test_canny_edges_full_size_equals_cv2():
    """1024x436 synthetic frames of the bench workload (+ noise so that long weak chains exist) vs cv2 itself."""
    ops = pkg("ops")
    img1, img2, _, _ = pkg("synth").make_pair(436, 1024, 0)
    rng = np.random.default_rng(3)
    noisy = np.clip(img2.astype(np.int32) + rng.integers(-40, 41, size=img2.shape), 0, 255).astype(np.uint8)
    total = 0
    for img, (lo, hi) in ((img1, (100, 200)), (img2, (20, 60)), (noisy, (100, 200)), (noisy, (40, 250))):
        bl = cv2.GaussianBlur(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), (3, 3), 0)
        want = np.array((255 - cv2.Canny(image=bl, threshold1=lo, threshold2=hi)) / 255, dtype="float32")
        got = ops.canny_edges(torch.from_numpy(img).cuda(), lo, hi).cpu().numpy()
        assert np.array_equal(got, want), (lo, hi, int((got != want).sum()))
        total += int((want == 0).sum())
    assert total > 10000


def test_canny_edges_bad_arguments():
    lib = pkg("_lib")
    L = lib.load()
    img = torch.zeros((8, 8, 3), dtype=torch.uint8, device="cuda")
    out = torch.zeros((8, 8), dtype=torch.float32, device="cuda")
    ws = torch.empty(16, dtype=torch.uint8, device="cuda")
    assert L.flowb200_canny_edges(img.data_ptr(), 8, 8, 100, 200, out.data_ptr(), ws.data_ptr(), 16, None) == lib.EWORKSPACE
    assert L.flowb200_canny_edges(None, 8, 8, 100, 200, out.data_ptr(), ws.data_ptr(), 16, None) == lib.EINVAL


def test_spremi_za_epic_script(tmp_path):
    """python spremiZaEpic.py <img1> <img2> <fwd.npy> <bwd.npy> <thr> canny (reference spremiZaEpic.py:1-27): the three
    EpicFlow inputs equal the reference's own outputs; the EpicFlow binary is absent here, so the run ends the way the
    reference's does without it (FileNotFoundError from subprocess.run, after all three files are written)."""
    import subprocess
    zc, ze = load_npz("consistency"), load_npz("edges")
    np.save(tmp_path / "f.npy", zc["real_fwd"])
    np.save(tmp_path / "b.npy", zc["real_bwd"])
    assert cv2.imwrite(str(tmp_path / "k1.png"), ze["boxes_img"])
    assert cv2.imwrite(str(tmp_path / "k2.png"), ze["noise_img"])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "spremiZaEpic.py"), "k1.png", "k2.png", "f.npy", "b.npy",
                        str(int(zc["real_thr"])), "canny"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "FileNotFoundError" in r.stderr, r.stderr[-2000:]
    assert np.array_equal(np.load(tmp_path / "sparse_field.npy"), zc["real_out"])
    with open(os.path.join(ROOT, "tests", "golden", "parovi_real.txt")) as f:
        assert (tmp_path / "parovi.txt").read_text() == f.read()
    e = np.fromfile(tmp_path / "ivice.bin", dtype=np.float32).reshape(ze["boxes_img"].shape[:2])
    assert np.array_equal(e == 0, ze["boxes_edges"] == 1)
    assert not (tmp_path / "epic.flo").exists()
