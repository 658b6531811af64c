"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the YOLOv8n body + Detect head (FP32, torch CPU).

Restates the network the reference runs through TensorRT
(/root/reference/src/yolo_engine.cpp:100-105, README.md:9-16,24-25).  The model file is not in the
reference tree (CMakeLists.txt:114), so the architecture is the published ultralytics YOLOv8n
(`yolov8.yaml` scale n, modules Conv/C2f/Bottleneck/SPPF/Detect/DFL) at nc = 14
(/root/reference/include/irmv_detection/armor.hpp:7) with BN already folded into conv weight+bias.
Parity status: "parity unpinned" -- the reference holds no golden vector for the network
(test/yolo_test.cpp:36 only bounds the count); the second opinion is OpenCV-DNN running the ONNX
export of this same module (see oracle/export_onnx.py, tests/test_oracle_cpu.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REG_MAX = 16
STRIDES = (8, 16, 32)


class Conv(nn.Module):
    """Conv2d(bias) [+ SiLU]: the BN-folded form of ultralytics `Conv`."""

    def __init__(self, c1, c2, k=1, s=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=True)
        self.act = act

    def forward(self, x):
        y = self.conv(x)
        return F.silu(y) if self.act else y


class Bottleneck(nn.Module):
    def __init__(self, c, shortcut):
        super().__init__()
        self.cv1 = Conv(c, c, 3, 1)
        self.cv2 = Conv(c, c, 3, 1)
        self.add = shortcut

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n, shortcut):
        super().__init__()
        self.c = c2 // 2
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, shortcut) for _ in range(n))
        self.cv2 = Conv((2 + n) * self.c, c2, 1, 1)

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        for m in self.m:
            y.append(m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2):
        super().__init__()
        self.cv1 = Conv(c1, c1 // 2, 1, 1)
        self.cv2 = Conv(c1 * 2, c2, 1, 1)

    def forward(self, x):
        x = self.cv1(x)
        y1 = F.max_pool2d(x, 5, 1, 2)
        y2 = F.max_pool2d(y1, 5, 1, 2)
        y3 = F.max_pool2d(y2, 5, 1, 2)
        return self.cv2(torch.cat((x, y1, y2, y3), 1))


class YoloV8n(nn.Module):
    """Module order matches irmv_detection_b200.weights.conv_specs()."""

    def __init__(self, nc=14, pose=False):
        super().__init__()
        self.nc = nc
        self.pose = pose
        self.m0 = Conv(3, 16, 3, 2)
        self.m1 = Conv(16, 32, 3, 2)
        self.m2 = C2f(32, 32, 1, True)
        self.m3 = Conv(32, 64, 3, 2)
        self.m4 = C2f(64, 64, 2, True)
        self.m5 = Conv(64, 128, 3, 2)
        self.m6 = C2f(128, 128, 2, True)
        self.m7 = Conv(128, 256, 3, 2)
        self.m8 = C2f(256, 256, 1, True)
        self.m9 = SPPF(256, 256)
        self.m12 = C2f(384, 128, 1, False)
        self.m15 = C2f(192, 64, 1, False)
        self.m16 = Conv(64, 64, 3, 2)
        self.m18 = C2f(192, 128, 1, False)
        self.m19 = Conv(128, 128, 3, 2)
        self.m21 = C2f(384, 256, 1, False)
        c3 = max(64, min(nc, 100))
        self.box = nn.ModuleList(
            nn.Sequential(Conv(ch, 64, 3), Conv(64, 64, 3), Conv(64, 4 * REG_MAX, 1, 1, act=False))
            for ch in (64, 128, 256))
        self.cls = nn.ModuleList(
            nn.Sequential(Conv(ch, c3, 3), Conv(c3, c3, 3), Conv(c3, nc, 1, 1, act=False))
            for ch in (64, 128, 256))
        if pose:            # ultralytics `Pose.cv4`, kpt_shape [4, 2]: c4 = max(ch[0] // 4, nk)
            self.kpt = nn.ModuleList(
                nn.Sequential(Conv(ch, 16, 3), Conv(16, 16, 3), Conv(16, 8, 1, 1, act=False))
                for ch in (64, 128, 256))

    def convs_in_order(self):
        out = [self.m0.conv, self.m1.conv]

        def c2f(m):
            r = [m.cv1.conv]
            for b in m.m:
                r += [b.cv1.conv, b.cv2.conv]
            return r + [m.cv2.conv]

        out += c2f(self.m2) + [self.m3.conv] + c2f(self.m4) + [self.m5.conv] + c2f(self.m6)
        out += [self.m7.conv] + c2f(self.m8) + [self.m9.cv1.conv, self.m9.cv2.conv]
        out += c2f(self.m12) + c2f(self.m15) + [self.m16.conv] + c2f(self.m18)
        out += [self.m19.conv] + c2f(self.m21)
        for i in range(3):
            out += [self.box[i][0].conv, self.box[i][1].conv, self.box[i][2].conv]
            out += [self.cls[i][0].conv, self.cls[i][1].conv, self.cls[i][2].conv]
        if self.pose:
            for i in range(3):
                out += [self.kpt[i][0].conv, self.kpt[i][1].conv, self.kpt[i][2].conv]
        return out

    def load_irmw(self, path):
        from irmv_detection_b200 import weights as W
        nc, tensors = W.load(path)
        assert nc == self.nc
        convs = self.convs_in_order()
        assert len(convs) == len(tensors)
        with torch.no_grad():
            for conv, (spec, w, b) in zip(convs, tensors):
                assert tuple(conv.weight.shape) == w.shape, spec.name
                conv.weight.copy_(torch.from_numpy(w))
                conv.bias.copy_(torch.from_numpy(b))
        return self

    def features(self, x, taps=None):
        """Returns per-scale raw head tensors [(box[B,64,H,W], cls[B,nc,H,W])]."""
        def tap(name, t):
            if taps is not None:
                taps[name] = t
            return t
        x0 = tap("m0", self.m0(x))
        x1 = tap("m1", self.m1(x0))
        x2 = tap("m2", self.m2(x1))
        x3 = tap("m3", self.m3(x2))
        x4 = tap("m4", self.m4(x3))
        x5 = tap("m5", self.m5(x4))
        x6 = tap("m6", self.m6(x5))
        x7 = tap("m7", self.m7(x6))
        x8 = tap("m8", self.m8(x7))
        x9 = tap("m9", self.m9(x8))
        u = F.interpolate(x9, scale_factor=2, mode="nearest")
        x12 = tap("m12", self.m12(torch.cat((u, x6), 1)))
        u = F.interpolate(x12, scale_factor=2, mode="nearest")
        x15 = tap("m15", self.m15(torch.cat((u, x4), 1)))
        x16 = tap("m16", self.m16(x15))
        x18 = tap("m18", self.m18(torch.cat((x16, x12), 1)))
        x19 = tap("m19", self.m19(x18))
        x21 = tap("m21", self.m21(torch.cat((x19, x9), 1)))
        outs = []
        for i, f in enumerate((x15, x18, x21)):
            o = (self.box[i](f), self.cls[i](f))
            outs.append(o + (self.kpt[i](f),) if self.pose else o)
        return outs

    def forward(self, x):
        """x f32[B,3,640,640] -> boxes xyxy f32[B,8400,4] (net px), scores f32[B,8400,nc]."""
        outs = self.features(x)
        if self.pose:
            return decode_heads(outs) + (decode_keypoints(outs),)
        return decode_heads(outs)


class DWConv(nn.Module):
    """Depthwise 3x3 conv with (BN-folded) bias, no activation (ShuffleNetV2 units)."""

    def __init__(self, c, s):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, s, 1, groups=c, bias=True)

    def forward(self, x):
        return self.conv(x)


def channel_shuffle(x, groups=2):
    b, c, h, w = x.shape
    return x.view(b, groups, c // groups, h, w).transpose(1, 2).reshape(b, c, h, w)


class ShuffleDown(nn.Module):
    """ShuffleNetV2 spatial down-sampling unit (Ma et al. 2018, fig. 3d), SiLU after the 1x1 convs."""

    def __init__(self, cin, cout):
        super().__init__()
        h = cout // 2
        self.b1_dw, self.b1_pw = DWConv(cin, 2), Conv(cin, h, 1, 1)
        self.b2_pw1, self.b2_dw, self.b2_pw2 = Conv(cin, h, 1, 1), DWConv(h, 2), Conv(h, h, 1, 1)

    def convs(self):
        return [self.b1_dw.conv, self.b1_pw.conv, self.b2_pw1.conv, self.b2_dw.conv, self.b2_pw2.conv]

    def forward(self, x):
        return channel_shuffle(torch.cat((self.b1_pw(self.b1_dw(x)), self.b2_pw2(self.b2_dw(self.b2_pw1(x)))), 1))


class ShuffleBasic(nn.Module):
    """ShuffleNetV2 basic unit (fig. 3c): channel split, right half 1x1 -> depthwise 3x3 -> 1x1, concat, shuffle."""

    def __init__(self, c):
        super().__init__()
        h = c // 2
        self.pw1, self.dw, self.pw2 = Conv(h, h, 1, 1), DWConv(h, 1), Conv(h, h, 1, 1)

    def convs(self):
        return [self.pw1.conv, self.dw.conv, self.pw2.conv]

    def forward(self, x):
        x1, x2 = x.chunk(2, 1)
        return channel_shuffle(torch.cat((x1, self.pw2(self.dw(self.pw1(x2)))), 1))


class ShuffleV2Kpt(YoloV8n):
    """The keypoint detector on a ShuffleNetV2-style backbone (irmv_detection_b200.weights.shuffle_conv_specs):
    stem + four down units (+ 0/1/3/1 basic units) produce P3/P4/P5-in, then SPPF, the YOLOv8n neck, Detect
    and the Pose branch.  Parity unpinned by the reference (the model is only named, README.md:12,16)."""

    def __init__(self, nc=14):
        super().__init__(nc, pose=True)
        from irmv_detection_b200 import weights as W
        for n in ("m1", "m2", "m3", "m4", "m5", "m6", "m7", "m8"):
            delattr(self, n)
        self.stages = nn.ModuleList()
        for name, cin, cout, units in W.shuffle_stage_plan():
            self.stages.append(nn.ModuleList([ShuffleDown(cin, cout)] + [ShuffleBasic(cout) for _ in range(units)]))

    def convs_in_order(self):
        out = [self.m0.conv]
        for st in self.stages:
            for u in st:
                out += u.convs()
        out += [self.m9.cv1.conv, self.m9.cv2.conv]

        def c2f(m):
            r = [m.cv1.conv]
            for b in m.m:
                r += [b.cv1.conv, b.cv2.conv]
            return r + [m.cv2.conv]

        out += c2f(self.m12) + c2f(self.m15) + [self.m16.conv] + c2f(self.m18) + [self.m19.conv] + c2f(self.m21)
        for i in range(3):
            out += [self.box[i][0].conv, self.box[i][1].conv, self.box[i][2].conv]
            out += [self.cls[i][0].conv, self.cls[i][1].conv, self.cls[i][2].conv]
        for i in range(3):
            out += [self.kpt[i][0].conv, self.kpt[i][1].conv, self.kpt[i][2].conv]
        return out

    def features(self, x, taps=None):
        def tap(name, t):
            if taps is not None:
                taps[name] = t
            return t
        x = tap("m0", self.m0(x))
        feats = []
        for i, st in enumerate(self.stages):
            for u in st:
                x = u(x)
            feats.append(tap(f"d{i + 1}", x))
        x4, x6 = feats[1], feats[2]
        x9 = tap("m9", self.m9(feats[3]))
        u = F.interpolate(x9, scale_factor=2, mode="nearest")
        x12 = tap("m12", self.m12(torch.cat((u, x6), 1)))
        u = F.interpolate(x12, scale_factor=2, mode="nearest")
        x15 = tap("m15", self.m15(torch.cat((u, x4), 1)))
        x16 = tap("m16", self.m16(x15))
        x18 = tap("m18", self.m18(torch.cat((x16, x12), 1)))
        x19 = tap("m19", self.m19(x18))
        x21 = tap("m21", self.m21(torch.cat((x19, x9), 1)))
        return [(self.box[i](f), self.cls[i](f), self.kpt[i](f)) for i, f in enumerate((x15, x18, x21))]


def make_anchors(sizes=(80, 40, 20)):
    pts, strides = [], []
    for hw, s in zip(sizes, STRIDES):
        ys, xs = np.meshgrid(np.arange(hw, dtype=np.float32) + 0.5,
                             np.arange(hw, dtype=np.float32) + 0.5, indexing="ij")
        pts.append(np.stack((xs.reshape(-1), ys.reshape(-1)), 1))
        strides.append(np.full((hw * hw,), s, np.float32))
    return np.concatenate(pts), np.concatenate(strides)


def decode_heads(outs):
    """DFL expectation + dist2bbox(xyxy) * stride, sigmoid class scores (ultralytics Detect/DFL).

    Anchor order: scale-major (80x80, 40x40, 20x20), row-major inside a scale.
    """
    B = outs[0][0].shape[0]
    box = torch.cat([o[0].reshape(B, 4 * REG_MAX, -1) for o in outs], 2)   # [B,64,A]
    cls = torch.cat([o[1].reshape(B, o[1].shape[1], -1) for o in outs], 2)  # [B,nc,A]
    A = box.shape[2]
    p = box.view(B, 4, REG_MAX, A).softmax(2)
    proj = torch.arange(REG_MAX, dtype=torch.float32).view(1, 1, REG_MAX, 1)
    dist = (p * proj).sum(2)                                               # [B,4,A] l,t,r,b
    anchors, strides = make_anchors(tuple(int(o[0].shape[2]) for o in outs))
    a = torch.from_numpy(anchors).t().unsqueeze(0)                         # [1,2,A]
    s = torch.from_numpy(strides).view(1, 1, A)
    x1y1 = (a - dist[:, :2]) * s
    x2y2 = (a + dist[:, 2:]) * s
    boxes = torch.cat((x1y1, x2y2), 1).permute(0, 2, 1).contiguous()       # [B,A,4]
    scores = cls.sigmoid().permute(0, 2, 1).contiguous()                   # [B,A,nc]
    return boxes, scores


def decode_keypoints(outs):
    """ultralytics `Pose.kpts_decode` for kpt_shape [4, 2]: (raw * 2 + (anchor - 0.5)) * stride, i.e.
    (raw * 2 + grid) * stride.  Returns f32[B, A, 4, 2] in network pixels; anchor order as decode_heads."""
    B = outs[0][2].shape[0]
    raw = torch.cat([o[2].reshape(B, 8, -1) for o in outs], 2)              # [B,8,A]
    A = raw.shape[2]
    anchors, strides = make_anchors(tuple(int(o[0].shape[2]) for o in outs))
    g = torch.from_numpy(anchors - 0.5).t().unsqueeze(0)                    # [1,2,A] grid x, y
    s = torch.from_numpy(strides).view(1, 1, 1, A)
    k = raw.view(B, 4, 2, A)
    return ((k * 2.0 + g.unsqueeze(1)) * s).permute(0, 3, 1, 2).contiguous()


def build(weights_path, nc=14):
    from irmv_detection_b200 import weights as W
    torch.manual_seed(0)
    if W.file_arch(weights_path) == W.ARCH_SHUFFLE_KPT:
        m = ShuffleV2Kpt(nc).eval()
    else:
        m = YoloV8n(nc, pose=len(W.load(weights_path)[1]) == 72).eval()
    m.load_irmw(weights_path)
    return m
