"""Optical-flow file IO in the reference's formats (tfoptflow/optflow.py:65-161), so flows predicted through this
library can be written next to, and compared with, the reference's own `.flo` files.

`.flo` (Middlebury): float32 tag 202021.25 ("PIEH"), int32 width, int32 height, then height x width x (u, v) float32,
row-major, little-endian.  `.pfm` (FlyingThings3D) and 16-bit `.png` (KITTI) are read-only, like in the reference."""
from __future__ import annotations

import os

import numpy as np

TAG_FLOAT = 202021.25  # tfoptflow/optflow.py:62


def flow_read(src_file: str) -> np.ndarray:
    """optflow.py:65-142 -> flow [h, w, 2] float32 (u = horizontal, v = vertical)."""
    if not os.path.exists(src_file):
        raise FileNotFoundError(src_file)
    low = src_file.lower()
    if low.endswith(".flo"):
        with open(src_file, "rb") as f:
            tag = float(np.fromfile(f, "<f4", count=1)[0])
            if tag != TAG_FLOAT:
                raise IOError(f"{src_file}: bad .flo tag {tag}")
            w = int(np.fromfile(f, "<i4", count=1)[0])
            h = int(np.fromfile(f, "<i4", count=1)[0])
            flow = np.fromfile(f, "<f4", count=h * w * 2)
            if flow.size != h * w * 2:
                raise IOError(f"{src_file}: truncated .flo ({flow.size} of {h * w * 2} values)")
            return flow.reshape(h, w, 2).astype(np.float32, copy=False)
    if low.endswith(".pfm"):
        with open(src_file, "rb") as f:
            tag = f.readline().rstrip().decode("utf-8")
            if tag != "PF":
                raise IOError(f"{src_file}: not a colour PFM")
            w, h = map(int, f.readline().rstrip().decode("utf-8").split(" "))
            scale = float(f.readline().rstrip().decode("utf-8"))
            flow = np.fromfile(f, "<f" if scale < 0 else ">f")
            return np.flipud(np.reshape(flow, (h, w, 3))[:, :, 0:2]).astype(np.float32)
    if low.endswith(".png"):
        import cv2  # KITTI 16-bit encoding, optflow.py:101-115
        raw = cv2.imread(src_file, -1)
        flow = raw[:, :, 2:0:-1].astype(np.float32)
        flow = (flow - 32768) / 64
        flow[np.abs(flow) < 1e-10] = 1e-10
        flow[raw[:, :, 0] == 0, :] = 0
        return flow
    raise IOError(f"{src_file}: unknown flow format")


def flow_write(flow: np.ndarray, dst_file: str) -> None:
    """optflow.py:145-161."""
    flow = np.asarray(flow)
    if flow.ndim != 3 or flow.shape[2] != 2:
        raise ValueError("flow must be [h, w, 2]")
    d = os.path.dirname(dst_file)
    if d:
        os.makedirs(d, exist_ok=True)
    if os.path.exists(dst_file):
        os.remove(dst_file)
    with open(dst_file, "wb") as f:
        np.array(TAG_FLOAT, dtype="<f4").tofile(f)
        h, w = flow.shape[:2]
        np.array(w, dtype="<u4").tofile(f)
        np.array(h, dtype="<u4").tofile(f)
        flow.astype("<f4").tofile(f)


def flow_mag_stats(flow: np.ndarray):
    """optflow.py:168-187: (min, avg, max) of the flow magnitude."""
    mag = np.sqrt(np.sum(np.square(np.asarray(flow, np.float64)), axis=-1))
    return float(mag.min()), float(mag.mean()), float(mag.max())
