"""TEST INFRASTRUCTURE ONLY -- derives the per-conv gain table baked into
irmv_detection_b200/weights.py::INIT_GAIN (LSUV-style, Mishkin & Matas 2016).

A BN-folded random-init network has no normalisation left, so plain fan-in scaling lets the
activation scale drift by orders of magnitude over 63 convs.  This script walks the convs in
execution order on a two-frame calibration batch (tests/golden/rm_test.jpg and a seeded uniform
frames), measures each conv's pre-activation std and rescales it to a target, and prints the
resulting gains.  Run:  python -m oracle.calibrate_init
"""
import math
import sys

import numpy as np
import torch


def main():
    import cv2
    from irmv_detection_b200 import weights as W
    from oracle import yolov8n_ref as Y, preprocess_ref as PR
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    arch = W.ARCH_SHUFFLE_KPT if (len(sys.argv) > 2 and sys.argv[2] == "shuffle") else W.ARCH_YOLOV8N
    specs = W.specs_for(arch)
    rng = np.random.default_rng(seed)
    tens = []
    for c in specs:
        cg = c.cin // c.groups
        w = rng.standard_normal((c.cout, cg, c.k, c.k)).astype(np.float32) / math.sqrt(cg * c.k * c.k)
        b = rng.standard_normal(c.cout).astype(np.float32) * 0.05
        tens.append((w, b))
    W.save("/tmp/_calib.irmw", tens, arch=arch)
    m = Y.build("/tmp/_calib.irmw")
    img = cv2.imread("tests/golden/rm_test.jpg")
    from irmv_detection_b200 import synth
    frames = [img] + list(synth.frames_from_base(img, 3, seed=1234))
    x = torch.from_numpy(np.stack([PR.preprocess(f)[0] for f in frames]))
    convs = m.convs_in_order()
    gains = []
    for c, conv in zip(specs, convs):
        target = 1.0
        top = c.name.split(".")[0]
        if top in ("m2", "m4", "m6", "m8") and ".m" in c.name and c.name.endswith("cv2"):
            target = 0.5                      # residual branch
        if c.name.startswith("m22.box") and c.name.endswith(".2"):
            target = 2.0
        if c.name.startswith("m22.cls") and c.name.endswith(".2"):
            target = 1.0
        if c.name.startswith("m22.kpt"):
            target = 0.1 if c.name.endswith(".2") else 1.0
        got = {}
        h = conv.register_forward_hook(lambda mod, i, o: got.__setitem__("s", float(o.std())))
        with torch.no_grad():
            m.features(x)
        h.remove()
        g = target / max(got["s"], 1e-6)
        with torch.no_grad():
            conv.weight.mul_(g)
        gains.append(g)
        print(f"{c.name:14s} std {got['s']:.4f} gain {g:.4f}", file=sys.stderr)
    print("INIT_GAIN = (")
    for i in range(0, len(gains), 8):
        print("    " + ", ".join(f"{g:.4g}" for g in gains[i:i + 8]) + ",")
    print(")")


if __name__ == "__main__":
    main()
