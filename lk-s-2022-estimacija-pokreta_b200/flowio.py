"""Flow file formats on either side of the hot path, vectorised restatements of the
reference's `visualization.py` codecs (host side, numpy + cv2 only; nothing here is timed):

  read_flo_file(path)                      visualization.py:9-29    Middlebury .flo -> float32 (h, w, 2) = (u, v)
  read_kitti_png(path)                     :37-53   FlowImage.readFlowFieldFromImage: 16-bit PNG -> (u, v, valid)
  write_flow_field(flow, path)             :75-82   FlowImage.writeFlowField (quirk Q9 below)
  load_flow(path)                          :98-124  FlowImage.ucitajFlow: .png / .npy / .flo -> float32 (h, w, 3)
  write_kitti_png(flow, path)              the KITTI devkit layout read_kitti_png expects (not in the reference)

Quirks reproduced: Q8 `.npy` with 3 channels is read as (dy, dx, valid) although `sparse_field.npy` holds
(dx, dy, valid) (:101-107); Q9 writeFlowField hands (u, v, 1) to cv2.imwrite in that channel order, so the file is
NOT the layout readFlowFieldFromImage reads back (it would see `valid` in the u channel), and the float -> uint16
conversion truncates.
"""
import os

import cv2
import numpy as np

FLO_MAGIC = np.float32(202021.25)


def read_flo_file(filename):
    with open(filename, "rb") as f:
        magic = np.fromfile(f, np.float32, count=1)
        if magic.size != 1 or magic[0] != FLO_MAGIC:
            print("Magic number incorrect. Invalid .flo file")
            return None
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        data = np.fromfile(f, np.float32, count=2 * w * h)
    return np.resize(data, (h, w, 2))


def write_flo_file(flow_uv, filename):
    """Inverse of read_flo_file (the reference only reads .flo)."""
    h, w = flow_uv.shape[:2]
    with open(filename, "wb") as f:
        np.array([FLO_MAGIC], np.float32).tofile(f)
        np.array([w, h], np.int32).tofile(f)
        np.ascontiguousarray(flow_uv[:, :, :2], np.float32).tofile(f)


def read_kitti_png(file_name):
    """(h, w, 3) float32 = (u, v, valid): pixels whose valid channel is 0 are (0, 0, 0)."""
    img = cv2.imread(file_name, -1)
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    valid = img[:, :, 2] > 0
    out = np.zeros(img.shape[:2] + (3,), dtype=np.float32)
    # the reference evaluates (uint16 - 32768.0) / 64.0 in float64 and stores float32
    out[:, :, 0] = np.where(valid, (img[:, :, 0].astype(np.float64) - 32768.0) / 64.0, 0.0)
    out[:, :, 1] = np.where(valid, (img[:, :, 1].astype(np.float64) - 32768.0) / 64.0, 0.0)
    out[:, :, 2] = valid
    return out


def write_flow_field(flow, path="flow_field.png"):
    """FlowImage.writeFlowField without the cv2.imshow: uint16 image [u*64+32768, v*64+32768, 1] at valid pixels,
    truncated to uint16, written with cv2.imwrite as is (Q9).  `np.float32 * 64.0 + 32768` is float32 arithmetic
    under NumPy >= 2 (weak Python scalars), which is what the golden vectors made here pin; NumPy 1.x evaluated it in
    float64, which differs in the last unit for values that round up at float32 spacing 2^-8."""
    flow = np.asarray(flow)
    image = np.zeros(flow.shape[:2] + (3,), dtype=np.uint16)
    valid = flow[:, :, 2] > 0.5
    u = flow[:, :, 0].astype(np.float32) * np.float32(64.0) + np.float32(32768)
    v = flow[:, :, 1].astype(np.float32) * np.float32(64.0) + np.float32(32768)
    image[:, :, 0] = np.where(valid, u, 0).astype(np.uint16)
    image[:, :, 1] = np.where(valid, v, 0).astype(np.uint16)
    image[:, :, 2] = valid
    cv2.imwrite(path, image)
    return image


def write_kitti_png(flow, path):
    """KITTI devkit layout (R, G, B) = (u*64+2^15, v*64+2^15, valid), i.e. what read_kitti_png reads back."""
    flow = np.asarray(flow)
    valid = flow[:, :, 2] > 0.5
    rgb = np.zeros(flow.shape[:2] + (3,), dtype=np.uint16)
    rgb[:, :, 0] = np.where(valid, np.clip(np.rint(flow[:, :, 0].astype(np.float64) * 64.0 + 32768.0), 0, 65535), 0)
    rgb[:, :, 1] = np.where(valid, np.clip(np.rint(flow[:, :, 1].astype(np.float64) * 64.0 + 32768.0), 0, 65535), 0)
    rgb[:, :, 2] = valid
    cv2.imwrite(path, cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR))


def load_flow(path):
    """FlowImage.ucitajFlow (:98-124): float32 (h, w, 3) = (u, v, valid), or None for an unknown extension."""
    ext = os.path.splitext(path)[1]
    if ext == ".png":
        return read_kitti_png(path)
    if ext == ".npy":
        flow = np.load(path)
        out = np.zeros(flow.shape[:2] + (3,), dtype=np.float32)
        out[:, :, 0] = flow[:, :, 1]
        out[:, :, 1] = flow[:, :, 0]
        # setValid(u, v, x): truthiness of the third channel (Q8: the channels of a 3-channel file are taken as dy, dx)
        out[:, :, 2] = (flow[:, :, 2] != 0) if flow.shape[2] == 3 else 1
        return out
    if ext == ".flo":
        flow = read_flo_file(path)
        out = np.zeros(flow.shape[:2] + (3,), dtype=np.float32)
        out[:, :, 0] = flow[:, :, 0]
        out[:, :, 1] = flow[:, :, 1]
        out[:, :, 2] = 1
        return out
    return None
