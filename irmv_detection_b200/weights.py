"""YOLOv8n (nc=14) convolution inventory, random-init generator and the `.irmw` weight file.

The reference loads a TensorRT engine sitting beside the ONNX path it is given
(/root/reference/src/yolo_engine.cpp:28-40, 137-151).  The armor checkpoints are not
available offline, so the new engine loads a flat weight file (`<onnx stem>.irmw`) that holds
the 63 BN-folded convolutions of the named architecture in a
fixed order.  Every stored value is exactly representable in FP16, so the FP32 oracle and the
FP16 tensor-core path start from identical parameters.

File layout (little endian):
    char[4] "IRMW" | u32 version(=1) | u32 nc | u32 n_convs (63, or 72 with the keypoint branch)
    per conv: u32 cin, cout, k, stride, act | f32 w[cout][cin][k][k] | f32 bias[cout]
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

MAGIC = b"IRMW"
VERSION = 1
NC = 14          # ArmorClass B1..RS, /root/reference/include/irmv_detection/armor.hpp:7
REG_MAX = 16
STRIDES = (8, 16, 32)
NET = 640        # hard-coded network size, /root/reference/src/yolo_engine.cpp:98-99,189-198


@dataclass(frozen=True)
class ConvSpec:
    name: str
    cin: int
    cout: int
    k: int
    stride: int
    act: int  # 1 = SiLU (Conv-BN-SiLU folded), 0 = plain conv2d with bias (Detect heads)


def _c2f(prefix: str, c1: int, c2: int, n: int) -> List[ConvSpec]:
    c = c2 // 2
    out = [ConvSpec(f"{prefix}.cv1", c1, 2 * c, 1, 1, 1)]
    for i in range(n):
        out.append(ConvSpec(f"{prefix}.m{i}.cv1", c, c, 3, 1, 1))
        out.append(ConvSpec(f"{prefix}.m{i}.cv2", c, c, 3, 1, 1))
    out.append(ConvSpec(f"{prefix}.cv2", (2 + n) * c, c2, 1, 1, 1))
    return out


NK = 8           # pose variant: kpt_shape = [4, 2], the four armor corners (x, y) per anchor


def conv_specs(nc: int = NC, pose: bool = False) -> List[ConvSpec]:
    """The 63 convolutions in execution order (ultralytics yolov8.yaml, scale n); with pose=True the
    nine convolutions of the keypoint branch (ultralytics `Pose.cv4`, kpt_shape [4, 2]) follow."""
    s: List[ConvSpec] = []
    s.append(ConvSpec("m0", 3, 16, 3, 2, 1))
    s.append(ConvSpec("m1", 16, 32, 3, 2, 1))
    s += _c2f("m2", 32, 32, 1)
    s.append(ConvSpec("m3", 32, 64, 3, 2, 1))
    s += _c2f("m4", 64, 64, 2)
    s.append(ConvSpec("m5", 64, 128, 3, 2, 1))
    s += _c2f("m6", 128, 128, 2)
    s.append(ConvSpec("m7", 128, 256, 3, 2, 1))
    s += _c2f("m8", 256, 256, 1)
    s.append(ConvSpec("m9.cv1", 256, 128, 1, 1, 1))
    s.append(ConvSpec("m9.cv2", 512, 256, 1, 1, 1))
    s += _c2f("m12", 384, 128, 1)
    s += _c2f("m15", 192, 64, 1)
    s.append(ConvSpec("m16", 64, 64, 3, 2, 1))
    s += _c2f("m18", 192, 128, 1)
    s.append(ConvSpec("m19", 128, 128, 3, 2, 1))
    s += _c2f("m21", 384, 256, 1)
    c2 = max(16, 64 // 4, 4 * REG_MAX)
    c3 = max(64, min(nc, 100))
    for i, ch in enumerate((64, 128, 256)):
        s.append(ConvSpec(f"m22.box{i}.0", ch, c2, 3, 1, 1))
        s.append(ConvSpec(f"m22.box{i}.1", c2, c2, 3, 1, 1))
        s.append(ConvSpec(f"m22.box{i}.2", c2, 4 * REG_MAX, 1, 1, 0))
        s.append(ConvSpec(f"m22.cls{i}.0", ch, c3, 3, 1, 1))
        s.append(ConvSpec(f"m22.cls{i}.1", c3, c3, 3, 1, 1))
        s.append(ConvSpec(f"m22.cls{i}.2", c3, nc, 1, 1, 0))
    assert len(s) == 63
    if pose:
        c4 = max(64 // 4, NK)
        for i, ch in enumerate((64, 128, 256)):
            s.append(ConvSpec(f"m22.kpt{i}.0", ch, c4, 3, 1, 1))
            s.append(ConvSpec(f"m22.kpt{i}.1", c4, c4, 3, 1, 1))
            s.append(ConvSpec(f"m22.kpt{i}.2", c4, NK, 1, 1, 0))
    return s


def total_flops(nc: int = NC) -> float:
    """2*MAC over the 63 convs at 640x640 (8.0956 GFLOP)."""
    hw = {}
    size = NET
    flops = 0.0
    # spatial size per conv follows the stride chain; recompute by walking the graph names
    res = {"m0": 320, "m1": 160, "m2": 160, "m3": 80, "m4": 80, "m5": 40, "m6": 40, "m7": 20,
           "m8": 20, "m9": 20, "m12": 40, "m15": 80, "m16": 40, "m18": 40, "m19": 20, "m21": 20}
    for c in conv_specs(nc):
        top = c.name.split(".")[0]
        if top == "m22":
            idx = int(c.name.split(".")[1][-1])
            r = (80, 40, 20)[idx]
        else:
            r = res[top]
        flops += 2.0 * r * r * c.cout * c.cin * c.k * c.k
    del hw, size
    return flops


# (box.2 gains are 0.2x the calibrated value: DFL logits of std ~0.4 keep the FP16-vs-FP32 box
# drift at stride 32 under the 0.5 px bar; with std 2 the heavy-tailed random logits reach +-54.)
# Per-conv gains (times 1/sqrt(fan_in)) from `python -m oracle.calibrate_init <seed>`: LSUV-style
# calibration that keeps every pre-activation at std ~1 on rm_test.jpg-like frames
# (synth.frames_from_base).  A BN-folded random-init net has no normalisation left, so the table
# is tied to the seed's random stream; it is baked in so the product generates weights with numpy
# only.  Seeds without a table reuse seed 0's gains (activation scale then drifts).
INIT_GAIN = {
    0: (
        13.79, 2.338, 1.84, 1.383, 1.154, 1.751, 1.385, 1.453,
        1.617, 0.8223, 1.589, 0.6623, 1.342, 1.352, 1.329, 1.602,
        0.7018, 1.385, 0.7448, 1.291, 1.495, 1.372, 1.452, 0.7605,
        1.424, 1.533, 0.3171, 1.517, 1.534, 2.132, 1.657, 1.609,
        1.538, 1.391, 1.453, 1.364, 1.679, 1.616, 1.398, 1.485,
        1.604, 1.495, 1.576, 1.457, 1.634, 1.312, 1.337, 0.595,
        1.267, 1.348, 1.306, 1.472, 1.548, 0.6308, 1.604, 1.683,
        1.401, 1.396, 1.564, 0.5298, 1.478, 1.828, 1.431,
    ),
    1: (
        15.52, 4.098, 1.296, 1.433, 0.9312, 1.669, 1.176, 1.504,
        2.025, 0.8036, 1.709, 0.8022, 1.553, 1.414, 1.763, 1.559,
        0.7482, 1.374, 0.7677, 1.38, 1.426, 1.494, 1.829, 0.7848,
        1.448, 1.38, 0.3566, 1.57, 1.762, 1.538, 1.795, 1.599,
        1.849, 1.546, 1.501, 1.493, 1.441, 1.124, 1.554, 1.358,
        1.58, 1.626, 1.619, 1.49, 1.655, 1.576, 1.671, 0.7328,
        1.445, 1.531, 1.47, 1.596, 1.276, 0.5678, 1.459, 1.502,
        1.718, 1.405, 1.744, 0.5846, 1.563, 1.503, 1.163,
    ),
}
CLS_BIAS = -8.0


def random_init(seed: int = 0, nc: int = NC, cls_bias: float = CLS_BIAS, pose: bool = False
                ) -> List[Tuple[np.ndarray, np.ndarray]]:
    """Seeded random-init, FP16-exact, activations kept O(1) through the SiLU chain.

    The class-logit bias plays the role of the ultralytics prior log(5/nc/(640/s)^2): with
    unit-std logits only a handful of (anchor, class) pairs clear the 0.25 score threshold
    (SURVEY.md section 7 "FP16 vs FP32 oracle drift").  seed 0 reproduces the calibration
    stream; other seeds reuse the same gains.
    """
    rng = np.random.default_rng(seed)
    out = []
    gains = INIT_GAIN.get(seed, INIT_GAIN[0])
    for i, c in enumerate(conv_specs(nc, pose)):
        fan_in = c.cin * c.k * c.k
        w = rng.standard_normal((c.cout, c.cin, c.k, c.k)).astype(np.float32) / np.float32(math.sqrt(fan_in))
        b = rng.standard_normal(c.cout).astype(np.float32) * np.float32(0.05)
        # (the keypoint branch comes after the 63 calibrated convs and draws from the stream last, so the
        # box/class network of a seed is the same with and without it; kpt.2 is kept small like box.2:
        # raw offsets of a few tenths of a cell)
        w = w * np.float32(gains[i] if i < len(gains) else ((0.05 if c.name.startswith("m22.kpt2") else 0.1)
                                                              if c.name.endswith(".2") else 1.5))
        if c.name.startswith("m22.cls") and c.name.endswith(".2"):
            b = b + np.float32(cls_bias)
        w = w.astype(np.float16).astype(np.float32)
        b = b.astype(np.float16).astype(np.float32)
        out.append((w, b))
    return out


def save(path: str, tensors: List[Tuple[np.ndarray, np.ndarray]], nc: int = NC) -> None:
    specs = conv_specs(nc, pose=len(tensors) == 72)
    assert len(specs) == len(tensors)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<III", VERSION, nc, len(specs)))
        for c, (w, b) in zip(specs, tensors):
            assert w.shape == (c.cout, c.cin, c.k, c.k) and b.shape == (c.cout,)
            f.write(struct.pack("<IIIII", c.cin, c.cout, c.k, c.stride, c.act))
            f.write(np.ascontiguousarray(w, dtype="<f4").tobytes())
            f.write(np.ascontiguousarray(b, dtype="<f4").tobytes())


def load(path: str) -> Tuple[int, List[Tuple[ConvSpec, np.ndarray, np.ndarray]]]:
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != MAGIC:
        raise ValueError(f"{path}: not an IRMW weight file")
    version, nc, n = struct.unpack_from("<III", data, 4)
    if version != VERSION:
        raise ValueError(f"{path}: unsupported version {version}")
    off = 16
    specs = conv_specs(nc, pose=n == 72)      # 63 convs: detector; 72: detector + keypoint branch
    if n != len(specs):
        raise ValueError(f"{path}: {n} convs, expected {len(specs)}")
    out = []
    for c in specs:
        cin, cout, k, stride, act = struct.unpack_from("<IIIII", data, off)
        off += 20
        if (cin, cout, k, stride, act) != (c.cin, c.cout, c.k, c.stride, c.act):
            raise ValueError(f"{path}: conv {c.name} header mismatch")
        nw = cout * cin * k * k
        w = np.frombuffer(data, "<f4", nw, off).reshape(cout, cin, k, k).copy()
        off += 4 * nw
        b = np.frombuffer(data, "<f4", cout, off).copy()
        off += 4 * cout
        out.append((c, w, b))
    if off != len(data):
        raise ValueError(f"{path}: trailing bytes")
    return nc, out


def write_random(path: str, seed: int = 0, nc: int = NC, pose: bool = False) -> str:
    save(path, random_init(seed, nc, pose=pose), nc)
    return path
