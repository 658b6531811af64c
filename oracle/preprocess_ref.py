"""TEST INFRASTRUCTURE ONLY -- CPU oracle for YoloEngine::preprocess.

Restates /root/reference/src/yolo_engine.cpp:179-200:
    K1 nppiMirror_8u_C3IR (NPP_BOTH_AXIS)  -> rot180 in place            (:182-184)
    K2 nppiResize_8u_C3R NPPI_INTER_LINEAR -> stretch to 640x640, u8     (:186-190)
    K3 nppiScale_8u32f_C3R [0,255]->[0,1]  -> x / 255                    (:192-194)
    K4 nppiCopy_32f_C3P3R                  -> packed HWC -> planar CHW   (:197-199)
The reference does NOT letterbox and does NOT swap channels on the device.

NPP's bilinear convention is not documented, so it was measured: oracle/npp_ref.cpp (the
reference's literal NPP chain, built into oracle/_ref/) was run on a B200 with NPP 12.4.1 and its
output for test/rm_test.jpg is pinned in tests/golden/npp_rm_golden.npz.  Result: nppiResize
LINEAR maps corner-aligned, src = dst * (src_size/dst_size) with NO half-pixel offset (so the
1280->640 axis simply picks every second pixel), and rounds to nearest.  `resize_bilinear_u8`
with half_pixel=False restates exactly that (bit-identical to the NPP output on the fixture);
half_pixel=True is the OpenCV/ultralytics convention kept as a non-reference mode.  Bayer input (config 4) is demosaiced first;
its oracle is cv2.cvtColor(COLOR_Bayer*2RGB) restated in `demosaic_bilinear`.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

NET = 640

# channel-order flags, mirrored by include/irmv_cabi.h
CH_PASSTHROUGH = 0   # reference behaviour: whatever is in the buffer goes to the net
CH_SWAP_RB = 1       # BGR <-> RGB
CH_BAYER_RGGB = 2
CH_BAYER_BGGR = 3
CH_BAYER_GRBG = 4
CH_BAYER_GBRG = 5


def rot180(img: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(img[::-1, ::-1])


def _axis_taps(n_dst: int, n_src: int, half_pixel: bool = False):
    """Source taps in FP32, the exact arithmetic the kernel uses.  half_pixel=False is NPP's
    measured corner-aligned map (reference-exact); True is the OpenCV-style pixel-centre map."""
    scale = np.float32(n_src) / np.float32(n_dst)
    d = np.arange(n_dst, dtype=np.float32)
    if half_pixel:
        s = (d + np.float32(0.5)) * scale - np.float32(0.5)
    else:
        s = d * scale
    i0 = np.floor(s)
    f = (s - i0).astype(np.float32)
    i0 = i0.astype(np.int64)
    i1 = i0 + 1
    i0c = np.clip(i0, 0, n_src - 1)
    i1c = np.clip(i1, 0, n_src - 1)
    return i0c, i1c, f


def resize_bilinear_f32(img: np.ndarray, out_h: int, out_w: int, half_pixel: bool = False) -> np.ndarray:
    """Bilinear stretch, FP32 result (no u8 rounding). img u8 [H,W,C]."""
    H, W, _ = img.shape
    y0, y1, fy = _axis_taps(out_h, H, half_pixel)
    x0, x1, fx = _axis_taps(out_w, W, half_pixel)
    p = img.astype(np.float32)
    fx_ = fx[None, :, None]
    fy_ = fy[:, None, None]
    one = np.float32(1.0)
    top = p[y0][:, x0] * (one - fx_) + p[y0][:, x1] * fx_
    bot = p[y1][:, x0] * (one - fx_) + p[y1][:, x1] * fx_
    return (top * (one - fy_) + bot * fy_).astype(np.float32)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int, half_pixel: bool = False) -> np.ndarray:
    v = resize_bilinear_f32(img, out_h, out_w, half_pixel)
    return np.clip(np.floor(v + np.float32(0.5)), 0, 255).astype(np.uint8)


def demosaic_bilinear(raw: np.ndarray, pattern: int) -> np.ndarray:
    """Bilinear demosaic of an 8-bit Bayer mosaic -> RGB u8 [H,W,3].

    Interior rule = cv2.cvtColor(COLOR_Bayer*2RGB): missing samples are the rounded mean of the
    2 or 4 nearest same-colour neighbours ((a+b+1)>>1, (a+b+c+d+2)>>2).  Borders use mirrored
    (reflect-101) neighbours, which keeps the colour phase; OpenCV instead replicates the
    adjacent interior pixel on the 1-px frame, so the two agree on [1:-1,1:-1] only.
    `pattern` names the 2x2 tile at the top-left: RGGB / BGGR / GRBG / GBRG.
    """
    H, W = raw.shape
    r = np.pad(raw.astype(np.int32), 1, mode="reflect")
    c = r[1:-1, 1:-1]
    n, s, w, e = r[:-2, 1:-1], r[2:, 1:-1], r[1:-1, :-2], r[1:-1, 2:]
    nw, ne, sw, se = r[:-2, :-2], r[:-2, 2:], r[2:, :-2], r[2:, 2:]
    cross = (n + s + w + e + 2) >> 2
    diag = (nw + ne + sw + se + 2) >> 2
    horiz = (w + e + 1) >> 1
    vert = (n + s + 1) >> 1
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    py, px = yy & 1, xx & 1
    # position of the red sample inside the 2x2 tile
    ry, rx = {CH_BAYER_RGGB: (0, 0), CH_BAYER_BGGR: (1, 1),
              CH_BAYER_GRBG: (0, 1), CH_BAYER_GBRG: (1, 0)}[pattern]
    is_r = (py == ry) & (px == rx)
    is_b = (py == (ry ^ 1)) & (px == (rx ^ 1))
    g_on_r_row = (py == ry) & (px != rx)      # green with red neighbours left/right
    g_on_b_row = (py != ry) & (px == rx)      # green with red neighbours above/below
    R = np.where(is_r, c, np.where(is_b, diag, np.where(g_on_r_row, horiz, vert)))
    B = np.where(is_b, c, np.where(is_r, diag, np.where(g_on_r_row, vert, horiz)))
    G = np.where(is_r | is_b, cross, c)
    return np.stack((R, G, B), 2).astype(np.uint8)


def mosaic_from_rgb(rgb: np.ndarray, pattern: int) -> np.ndarray:
    """Synthesise a Bayer mosaic from an RGB image (bench/test input generator)."""
    H, W, _ = rgb.shape
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    py, px = yy & 1, xx & 1
    ry, rx = {CH_BAYER_RGGB: (0, 0), CH_BAYER_BGGR: (1, 1),
              CH_BAYER_GRBG: (0, 1), CH_BAYER_GBRG: (1, 0)}[pattern]
    is_r = (py == ry) & (px == rx)
    is_b = (py == (ry ^ 1)) & (px == (rx ^ 1))
    return np.where(is_r, rgb[..., 0], np.where(is_b, rgb[..., 2], rgb[..., 1])).astype(np.uint8)


def to_rgb(src: np.ndarray, chan: int) -> np.ndarray:
    """Source buffer -> 3-channel u8 image in the order fed to the net."""
    if chan == CH_PASSTHROUGH:
        return src
    if chan == CH_SWAP_RB:
        return np.ascontiguousarray(src[..., ::-1])
    return demosaic_bilinear(src, chan)


def preprocess(src: np.ndarray, chan: int = CH_PASSTHROUGH, rotate: bool = True,
               quantize_u8: bool = True, half_pixel: bool = False):
    """Returns (input f32[3,640,640], rotated u8 image [H,W,3]) as the reference's graph leaves them.

    quantize_u8=True keeps the reference's 8-bit intermediate (K2 writes u8 before K3 scales it);
    False is the fused-precision mode (lerp kept in FP32).
    """
    img = to_rgb(src, chan)
    if rotate:
        img = rot180(img)
    if quantize_u8:
        r = resize_bilinear_u8(img, NET, NET, half_pixel).astype(np.float32)
    else:
        r = resize_bilinear_f32(img, NET, NET, half_pixel)
    x = (r / np.float32(255.0)).astype(np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1)), img


def preprocess_fp16(src, chan=CH_PASSTHROUGH, rotate=True, quantize_u8=True, half_pixel=False):
    """What the CUDA kernel stores: the FP32 result rounded to FP16 (round-to-nearest-even)."""
    x, img = preprocess(src, chan, rotate, quantize_u8, half_pixel)
    return x.astype(np.float16), img


# ---- LETTERBOX mode (not what the reference does: it stretches, src/yolo_engine.cpp:186-190; this is
# the form BASELINE.json's north_star / configs[0] name).  Restates ultralytics' LetterBox
# (data/augment.py, the preprocessing every YOLOv8 export expects): r = min(640/h, 640/w),
# new_unpad = round(size * r), cv2.resize(INTER_LINEAR) = pixel-centre bilinear, centred with
# top/left = round(d - 0.1), border value 114.  Parity unpinned by the reference.
LETTERBOX_PAD = 114


def letterbox_geometry(src_w: int, src_h: int):
    r = min(NET / src_w, NET / src_h)
    new_w, new_h = min(int(src_w * r + 0.5), NET), min(int(src_h * r + 0.5), NET)
    pad_x = max(int((NET - new_w) / 2.0 - 0.1 + 0.5), 0)
    pad_y = max(int((NET - new_h) / 2.0 - 0.1 + 0.5), 0)
    return pad_x, pad_y, new_w, new_h


def preprocess_letterbox(src: np.ndarray, chan: int = CH_PASSTHROUGH, rotate: bool = True, quantize_u8: bool = True):
    """(input f32[3,640,640], geometry) for resize_mode LETTERBOX."""
    img = to_rgb(src, chan)
    if rotate:
        img = rot180(img)
    H, W, _ = img.shape
    pad_x, pad_y, new_w, new_h = letterbox_geometry(W, H)
    if quantize_u8:
        r = resize_bilinear_u8(img, new_h, new_w, True).astype(np.float32)
        pad = np.float32(LETTERBOX_PAD)
    else:
        r = resize_bilinear_f32(img, new_h, new_w, True)
        pad = np.float32(LETTERBOX_PAD)
    out = np.full((NET, NET, 3), pad, np.float32)
    out[pad_y:pad_y + new_h, pad_x:pad_x + new_w] = r
    x = out / np.float32(255.0)
    return np.ascontiguousarray(x.transpose(2, 0, 1)), (pad_x, pad_y, new_w, new_h)


def unletterbox_boxes(boxes: np.ndarray, src_w: int, src_h: int) -> np.ndarray:
    """Network-pixel xyxy -> source pixels (the inverse the engine's parse_output applies)."""
    pad_x, pad_y, new_w, new_h = letterbox_geometry(src_w, src_h)
    sx, sy = np.float32(src_w) / np.float32(new_w), np.float32(src_h) / np.float32(new_h)
    b = np.asarray(boxes, np.float32).copy()
    b[:, [0, 2]] = (b[:, [0, 2]] - np.float32(pad_x)) * sx
    b[:, [1, 3]] = (b[:, [1, 3]] - np.float32(pad_y)) * sy
    return b


def preprocess_cv2(src_bgr: np.ndarray):
    """The CPU-baseline form named in BASELINE.md section 3 (cv2.flip + cv2.resize + /255 + CHW)."""
    import cv2
    img = cv2.flip(src_bgr, -1)
    r = cv2.resize(img, (NET, NET), interpolation=cv2.INTER_LINEAR)
    return np.ascontiguousarray((r.astype(np.float32) / 255.0).transpose(2, 0, 1)), img
