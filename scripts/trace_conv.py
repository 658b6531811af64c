"""Per-tile timeline of CTA 0 for selected GEMMs (debug)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import irmv_detection_b200 as irmv
    from irmv_detection_b200 import weights, _lib
    import bench
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    ops = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 8, 46, 47, 48]
    w = "/tmp/trace_seed0.irmw"
    weights.write_random(w, 0)
    dev = torch.device("cuda", 0)
    frames = bench.make_bayer_frames_device(n, 0, dev)
    torch.cuda.synchronize()
    eng = irmv.YoloEngine(w, (1280, 1024), chan_order=irmv.CH_BAYER_RGGB, max_batch=n, sub_batch=n, num_lanes=1, use_graph=False)
    for _ in range(int(os.environ.get("WARM", "2"))):
        eng.enqueue_batch_device(frames.data_ptr(), n)
        eng.sync()
    lib = _lib.lib()
    cap = 64
    for op in ops:
        buf = np.zeros((cap + 16, 8), np.int64)
        ms = C.c_float(0)
        for _ in range(2):
            k = lib.irmv_engine_trace_conv(eng._h, op, n, buf.ctypes.data, cap, C.byref(ms))
        t = buf[:k].astype(np.float64)
        if os.environ.get("RASTER"):
            t = t[(t != 0).all(axis=1)]
            k = len(t)
        t0 = t[0, 0]
        print(f"== op {op}: kernel {ms.value*1e3:.1f} us, tiles/CTA {k}")
        names = ["start", "pre-wait", "post-wait", "issued", "mma-first", "mma-last", "epi-start", "epi-end"]
        names = ["ld-start", "ld-gotslot", "ld-issued", "mma-wait", "mma-go", "mma-issued", "epi-start", "epi-end"] if os.environ.get("RASTER") else names
        for i in list(range(min(k, 6))) + list(range(max(6, k - 3), k)):
            print(f"  tile {i:3d} " + " ".join(f"{nm}={int(v - t0):7d}" for nm, v in zip(names, t[i])))
        if os.environ.get("RASTER"):
            for nm, row in (("CTA0", buf[cap + 15]), ("CTAlast", buf[cap + 14])):
                print(f"  {nm}: entry->setup {row[1]-row[0]} cyc, entry->exit {row[2]-row[0]} cyc = {row[4]-row[3]} ns, first ld-start at +{int(buf[0,0]-row[0])} cyc")
            continue
        kb = buf[cap:cap + 16].astype(np.float64)
        names2 = ["p:pre-empty", "p:post-empty", "p:committed", "p:waited", "p:arrived", "m:pre-full", "m:post-full", "m:committed"]
        for j in range(16):
            if kb[j].any():
                print(f"  tile3 kb {j:2d} " + " ".join(f"{nm}={int(v - t0) if v else -1:7d}" for nm, v in zip(names2, kb[j])))
        print(f"  CTA0 span {int(t[:, 7].max() - t0)} cycles over {k} tiles; kernel at 1.965 GHz would be {(t[:, 7].max() - t0) / 1965:.1f} us")
        if k > 4:
            d = np.diff(t[2:k, 7])
            print(f"  steady cycles/tile (epi-end to epi-end): median {np.median(d):.0f} mean {d.mean():.0f}")
    eng.close()


if __name__ == "__main__":
    main()
