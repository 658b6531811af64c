"""Per-layer CUDA-event times of one eager replay: `python scripts/profile_ops.py [frames]`."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    import torch
    import irmv_detection_b200 as irmv
    from irmv_detection_b200 import weights, _lib
    from analyze_launches import name_ops
    import bench
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    w = "/tmp/profile_seed0.irmw"
    weights.write_random(w, 0)
    frames = bench.make_bayer_frames_device(n, 0, torch.device("cuda", 0))
    torch.cuda.synchronize()
    eng = irmv.YoloEngine(w, (1280, 1024), chan_order=irmv.CH_BAYER_RGGB, max_batch=n, sub_batch=n, num_lanes=1, use_graph=False)
    for _ in range(3):
        eng.enqueue_batch_device(frames.data_ptr(), n)
        eng.sync()
    lib = _lib.lib()
    runs = []
    for _ in range(5):
        ms = np.zeros(80, np.float32)
        k = lib.irmv_engine_profile_ops(eng._h, C.c_void_p(frames.data_ptr()), n, ms.ctypes.data, 80)
        runs.append(ms[:k].copy())
    ms = np.median(np.stack(runs), axis=0) * 1e3
    L = name_ops(eng.describe_ops())      # conv0 lives in the stem kernel
    assert len(L) == len(ms), (len(L), len(ms))
    tot = ms.sum()
    print(f"network stage: {tot:.1f} us for {n} frames = {tot / n:.2f} us/frame ({8.0956e9 * n / tot / 1e6:.0f} TFLOP/s)")
    print(f"{'layer':16s} {'hw':>4s} {'cin':>4s} {'cout':>4s} k s {'us':>8s} {'share':>6s} {'TFLOP/s':>8s} {'GB/s(in+out)':>12s}")
    for (name, hw, cin, cout, k, s, flf), v in zip(L, ms):
        if name == "POOL":
            print(f"{name:16s} {'':18s} {v:8.1f} {100 * v / tot:5.1f}%")
            continue
        M = n * hw * hw
        fl = flf * n
        io = (n * (hw * s) ** 2 * cin + M * cout) * 2.0
        print(f"{name:16s} {hw:4d} {cin:4d} {cout:4d} {k} {s} {v:8.1f} {100 * v / tot:5.1f}% {fl / v / 1e6:8.1f} {io / v / 1e3:12.1f}")
    eng.close()


if __name__ == "__main__":
    main()
