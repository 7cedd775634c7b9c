"""Flow file codecs against outputs of the reference's own visualization.py code (tests/golden/
flowio.npz, made by tests/golden/make_golden_flowio.py)."""
import os

import cv2
import numpy as np

from helpers import pkg

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flowio.npz"))


def test_write_flow_field_matches_reference_bytes(tmp_path):
    fio = pkg("flowio")
    img = fio.write_flow_field(GOLD["flow"], str(tmp_path / "flow_field.png"))
    back = cv2.imread(str(tmp_path / "flow_field.png"), -1)
    assert np.array_equal(back, GOLD["written_png"]) and np.array_equal(img, GOLD["written_png"])


def test_readers_match_reference(tmp_path):
    fio = pkg("flowio")
    cv2.imwrite(str(tmp_path / "gt.png"), GOLD["kitti_png_bgr"])
    assert np.array_equal(fio.load_flow(str(tmp_path / "gt.png")), GOLD["read_png"])
    flow = GOLD["flow"]
    np.save(tmp_path / "f2.npy", np.stack([flow[:, :, 1], flow[:, :, 0]], -1).astype(np.float64))
    np.save(tmp_path / "f3.npy", flow)
    assert np.array_equal(fio.load_flow(str(tmp_path / "f2.npy")), GOLD["read_npy2"])
    assert np.array_equal(fio.load_flow(str(tmp_path / "f3.npy")), GOLD["read_npy3"])          # quirk Q8
    open(tmp_path / "f.flo", "wb").write(GOLD["flo_bytes"].tobytes())
    assert np.array_equal(fio.read_flo_file(str(tmp_path / "f.flo")), GOLD["read_flo_raw"])
    assert np.array_equal(fio.load_flow(str(tmp_path / "f.flo")), GOLD["read_flo"])
    assert fio.load_flow(str(tmp_path / "x.txt")) is None


def test_kitti_png_and_flo_round_trips(tmp_path):
    fio = pkg("flowio")
    flow = GOLD["flow"].copy()
    flow[:, :, :2] = np.round(flow[:, :, :2] * 64.0) / 64.0          # representable in 1/64 px
    flow[flow[:, :, 2] < 0.5] = 0
    fio.write_kitti_png(flow, str(tmp_path / "k.png"))
    assert np.array_equal(fio.read_kitti_png(str(tmp_path / "k.png")), flow)
    fio.write_flo_file(flow, str(tmp_path / "k.flo"))
    assert np.array_equal(fio.read_flo_file(str(tmp_path / "k.flo")), flow[:, :, :2])
