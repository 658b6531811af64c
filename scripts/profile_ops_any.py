"""Per-kernel CUDA-event times of one eager replay for any weight file / architecture:
`python scripts/profile_ops_any.py [frames] [yolov8n|yolov8n-pose|shufflenetv2-pose] [unfused]`."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import irmv_detection_b200 as irmv
    from irmv_detection_b200 import weights, _lib, synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    arch = sys.argv[2] if len(sys.argv) > 2 else "shufflenetv2-pose"
    w = f"/tmp/profile_{arch}.irmw"
    weights.write_random(w, 0, pose=(arch == "yolov8n-pose"), arch=arch)
    frames = torch.from_numpy(synth.frames_from_base(synth.load_base(), n, seed=2)).cuda()
    fuse = not (len(sys.argv) > 3 and sys.argv[3] == "unfused")
    eng = irmv.YoloEngine(w, (1280, 1024), max_batch=n, sub_batch=n, num_lanes=1, use_graph=False, fuse_units=fuse)
    for _ in range(3):
        eng.enqueue_batch_device(frames.data_ptr(), n)
        eng.sync()
    lib = _lib.lib()
    runs = []
    for _ in range(5):
        ms = np.zeros(128, np.float32)
        k = lib.irmv_engine_profile_ops(eng._h, C.c_void_p(frames.data_ptr()), n, ms.ctypes.data, 128)
        runs.append(ms[:k].copy())
    ms = np.median(np.stack(runs), axis=0) * 1e3
    ops = eng.describe_ops()
    assert len(ops) == len(ms), (len(ops), len(ms))
    print(f"{arch}: network stage {ms.sum():.1f} us for {n} frames, {len(ops)} kernels")
    for o, v in zip(ops, ms):
        if o["kind"] == "conv":
            io = n * ((o["hw"] * o["s"]) ** 2 * o["cin"] + o["hw"] ** 2 * max(o["cout"], o["tail_cout"])) * 2.0
            print(f"conv k{o['k']} s{o['s']} {o['cin']:4d}->{o['cout']:4d} hw {o['hw']:3d} {'raster' if o['raster'] else 'gather'} tail {o['tail_cout']:3d} {v:8.1f} us {io / v / 1e3:8.1f} GB/s(in+out)")
        elif o["kind"] == "dw":
            io = n * ((o["hw"] * o["s"]) ** 2 + o["hw"] ** 2) * o["c"] * 2.0
            print(f"dw   k3 s{o['s']} {o['c']:4d}         hw {o['hw']:3d}                 {v:8.1f} us {io / v / 1e3:8.1f} GB/s(in+out)")
        elif o["kind"] == "unit":
            s_ = 2 if o["down"] else 1
            io = n * ((o["hw"] * s_) ** 2 * o["cin"] * (1 if o["down"] else 2) + o["hw"] ** 2 * 2 * o["h"]) * 2.0
            print(f"unit {'down ' if o['down'] else 'basic'} {o['cin']:4d}->{2 * o['h']:4d} hw {o['hw']:3d} rows/cta {o['rows_per_cta']} {v:8.1f} us {io / v / 1e3:8.1f} GB/s(in+out)")
        else:
            print(f"pool {v:8.1f} us")
    eng.close()


if __name__ == "__main__":
    main()
