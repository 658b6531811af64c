"""Synthetic camera frames for tests and benches (there is no camera and no dataset offline).

SURVEY.md section 8d: frames are derived from the reference's only fixture, test/rm_test.jpg
(a dark 1280x1024 arena frame, copied to tests/golden/rm_test.jpg), with per-frame seeded
shifts, gain and sensor noise so that detections differ from frame to frame.
"""
from __future__ import annotations

import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BASE_IMAGE = os.path.join(os.path.dirname(_HERE), "tests", "golden", "rm_test.jpg")


def load_base(path: str = BASE_IMAGE) -> np.ndarray:
    """BGR u8 [1024,1280,3] as cv::imread returns it (/root/reference/test/yolo_test.cpp:24,63)."""
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise FileNotFoundError(path)
    return img


def frames_from_base(base: np.ndarray, n: int, seed: int = 0) -> np.ndarray:
    """n frames u8 [n,H,W,3]: circular shift (|dx|,|dy| <= 96), gain 0.9..1.2, N(0,3) noise."""
    rng = np.random.default_rng(seed)
    out = np.empty((n,) + base.shape, np.uint8)
    for i in range(n):
        dx, dy = rng.integers(-96, 97, 2)
        gain = rng.uniform(0.9, 1.2)
        f = np.roll(base, (int(dy), int(dx)), (0, 1)).astype(np.float32) * np.float32(gain)
        f += rng.normal(0.0, 3.0, base.shape).astype(np.float32)
        out[i] = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    return out


def bayer_from_rgb(rgb: np.ndarray, pattern: str = "RGGB") -> np.ndarray:
    """Sample an RGB frame [.., H, W, 3] onto an 8-bit Bayer mosaic [.., H, W]."""
    ry, rx = {"RGGB": (0, 0), "BGGR": (1, 1), "GRBG": (0, 1), "GBRG": (1, 0)}[pattern]
    out = rgb[..., 1].copy()
    out[..., ry::2, rx::2] = rgb[..., ry::2, rx::2, 0]
    out[..., (ry ^ 1)::2, (rx ^ 1)::2] = rgb[..., (ry ^ 1)::2, (rx ^ 1)::2, 2]
    return out
