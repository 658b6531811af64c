"""Smallest end-to-end run for compute-sanitizer: two frames through both architectures with the armor and
pose stages on (`compute-sanitizer --tool memcheck python scripts/sanitize_small.py`)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import irmv_detection_b200 as irmv
    from irmv_detection_b200 import synth, weights
    K, D = synth.K_CAMERA, synth.D_CAMERA
    base = synth.load_base()
    fr = synth.frames_from_base(base, 2, seed=1)
    raw = synth.bayer_from_rgb(fr[..., ::-1], "RGGB")
    for arch, pose in (("yolov8n", False), ("shufflenetv2-pose", False)):
        w = f"/tmp/san_{arch}.irmw"
        weights.write_random(w, 0, pose=pose, arch=arch)
        for chan, frames in ((irmv.CH_BAYER_RGGB, raw), (irmv.CH_PASSTHROUGH, fr)):
            eng = irmv.YoloEngine(w, (1280, 1024), chan_order=chan, max_batch=2, sub_batch=2, use_graph=False)
            eng.enable_armors()
            eng.enable_pnp(K, D, (0.5, 480 / 1024))
            d = eng.detect_batch(frames)
            eng.fetch_armor_poses(2)
            print(arch, chan, [len(x) for x in d])
            eng.close()
    s = irmv.PnPSolver(K, D)
    s.solve_batch(synth.armor_quads(64, 0), extended=True)
    s.close()


if __name__ == "__main__":
    main()
