"""Light-bar / armor extraction: the CUDA stage vs the reference's
per-box OpenCV chain (cvtColor, threshold, findContours, minAreaRect) on the host cores.
`python scripts/bench_armors.py [frames] [boxes per frame]` prints one JSON line."""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(n: int = 64, per: int = 10) -> dict:
    import torch
    import irmv_detection_b200 as irmv
    from irmv_detection_b200 import engine as E
    from irmv_detection_b200 import synth
    from oracle import armor_ref as A                          # CPU-baseline leg: the reference's cv2 chain
    scenes, boxes = [], np.zeros((n, 100), irmv.BBOX_DTYPE)
    for f in range(n):
        img, b, s, c = synth.armor_scene(per, 1000 + f)
        scenes.append(img)
        boxes["xyxy"][f, :per] = b; boxes["score"][f, :per] = s; boxes["class_id"][f, :per] = c
    counts = np.full(n, per, np.int32)
    frames = np.stack([np.ascontiguousarray(s[::-1, ::-1]) for s in scenes])              # camera view
    dev = torch.from_numpy(frames).cuda()
    torch.cuda.synchronize()
    ms = []
    for it in range(8):
        out = E.extract_armors_ptr(dev.data_ptr(), True, n, 1280, 1024, boxes, counts)
        if it >= 3:
            ms.append(E.extract_armors_last_device_ms())
    k = float(np.median(ms))
    prof = None
    if os.environ.get("IRMV_ARMOR_PROF"):
        import ctypes as C
        from irmv_detection_b200 import _lib
        buf = (C.c_uint64 * 8)()
        _lib.lib().irmv_extract_armors_last_profile(C.byref(buf))
        v = [int(x) for x in buf]
        prof = {"rois": v[4], "cycles_per_roi": {"bitmap": v[0] / max(v[4], 1), "flood": v[1] / max(v[4], 1),
                                                 "walks_lights": v[2] / max(v[4], 1), "armor": v[3] / max(v[4], 1)},
                "flood_rounds_per_roi": v[5] / max(v[4], 1),
                "warp_cycles_per_roi": {"recording_walks": v[6] / max(v[4], 1), "hull_rect_light": v[7] / max(v[4], 1)}}
    # ROI bytes the stage has to read (packed u8x3) + one 56-byte armor per detection
    roi_px = 0
    for f in range(n):
        for b in boxes["xyxy"][f, :per]:
            r = A.roi_of(b, 1280, 1024)
            roi_px += r[2] * r[3] if r else 0
    algo_bytes = roi_px * 3 + n * per * 56
    cores = os.cpu_count() or 1

    def cpu(f):
        return A.extract_armors_cv2(scenes[f], boxes["xyxy"][f, :per], boxes["score"][f, :per], boxes["class_id"][f, :per])
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        ref = list(ex.map(cpu, range(n)))
    cpu_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    cpu(0)
    cpu1_s = time.perf_counter() - t0
    same = all([a.bbox_index for a in ref[f]] == np.nonzero(out[f]["valid"])[0].tolist() for f in range(n))
    return {
        "workload": f"{n} frames 1280x1024x3, {per} detections each (seeded light-bar scenes)",
        "gpu_kernel_ms": k, "gpu_rois_per_s": n * per / (k * 1e-3), "gpu_us_per_frame": k * 1e3 / n,
        "algorithmic_bytes": algo_bytes, "hbm_gbs_achieved": algo_bytes / (k * 1e-3) / 1e9,
        "cpu_rois_per_s_all_cores": n * per / cpu_s, "cpu_cores": cores, "cpu_ms_per_frame_one_core": cpu1_s * 1e3,
        "speedup_vs_cpu_all_cores": (n * per / (k * 1e-3)) / (n * per / cpu_s),
        "phase_profile": prof,
        "armors": int(sum(int(o["valid"].sum()) for o in out)), "valid_sets_equal_cv2": bool(same),
    }


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    print(json.dumps(measure(n, per)))


if __name__ == "__main__":
    main()
