"""Runs the reference's NPP chain (oracle/_ref/npp_ref) on a GPU box and compares it with the
oracle restatement and the CUDA kernel; stores the NPP outputs as u8 (x*255) under gpurun_out/."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def main():
    import irmv_detection_b200 as irmv
    from irmv_detection_b200 import synth
    from oracle import preprocess_ref as PR
    exe = os.path.join(ROOT, "oracle", "_ref", "npp_ref")
    base = synth.load_base()
    rnd = np.random.default_rng(0).integers(0, 256, base.shape, dtype=np.uint8)
    grad = np.zeros_like(base)
    grad[..., 0] = (np.arange(1280)[None, :] % 256)
    grad[..., 1] = (np.arange(1024)[:, None] % 256)
    grad[..., 2] = ((np.arange(1280)[None, :] + np.arange(1024)[:, None]) % 256)
    for name, img in (("rm", base), ("rnd", rnd), ("grad", grad)):
        raw = f"/tmp/{name}.raw"
        img.tofile(raw)
        r = subprocess.run([exe, raw, "1280", "1024", f"/tmp/{name}.f32", f"/tmp/{name}_rot.raw"], capture_output=True, text=True)
        print(name, r.returncode, r.stdout.strip(), r.stderr.strip())
        if r.returncode:
            continue
        npp = np.fromfile(f"/tmp/{name}.f32", np.float32).reshape(3, 640, 640)
        rot = np.fromfile(f"/tmp/{name}_rot.raw", np.uint8).reshape(base.shape)
        u8 = np.rint(npp * 255.0)
        print("  npp output is k/255 exactly:", bool(np.array_equal((u8 / 255.0).astype(np.float32), npp)),
              "max |x*255 - round|", float(np.abs(npp * 255.0 - u8).max()))
        np.save(os.path.join(OUT, f"npp_{name}_u8.npy"), u8.astype(np.uint8))
        ref, ref_rot = PR.preprocess(img)
        print("  rot180 equal:", bool(np.array_equal(rot, ref_rot)))
        d = np.rint(ref * 255.0) - u8
        print("  oracle(half-pixel, round-half-up) vs NPP: max |d| (8-bit steps)", float(np.abs(d).max()),
              "frac != 0", float((d != 0).mean()), "hist", np.bincount((d + 3).astype(int).ravel(), minlength=7).tolist())
        ours = irmv.preprocess(img[None])[0, :, :, :3].astype(np.float32).transpose(2, 0, 1)
        d2 = np.rint(ours * 255.0) - u8
        print("  CUDA kernel vs NPP: max |d|", float(np.abs(d2).max()), "frac != 0", float((d2 != 0).mean()))
        # alternative conventions
        import cv2
        rimg = ref_rot
        for label, arr in (
            ("cv2 INTER_LINEAR", cv2.resize(rimg, (640, 640), interpolation=cv2.INTER_LINEAR)),
            ("cv2 INTER_AREA", cv2.resize(rimg, (640, 640), interpolation=cv2.INTER_AREA)),
            ("cv2 INTER_NEAREST", cv2.resize(rimg, (640, 640), interpolation=cv2.INTER_NEAREST)),
            ("corner-aligned sx=dx*scale", corner(rimg)),
            ("half-pixel floor", np.floor(PR.resize_bilinear_f32(rimg, 640, 640))),
            ("half-pixel rint", np.rint(PR.resize_bilinear_f32(rimg, 640, 640))),
        ):
            dd = arr.astype(np.float32).transpose(2, 0, 1) - u8
            print(f"    {label:28s} max |d| {np.abs(dd).max():5.1f} frac != 0 {(dd != 0).mean():.5f}")


def corner(img):
    H, W, _ = img.shape
    sy = np.arange(640, dtype=np.float32) * np.float32(H / 640)
    sx = np.arange(640, dtype=np.float32) * np.float32(W / 640)
    y0 = np.floor(sy).astype(int); x0 = np.floor(sx).astype(int)
    fy = (sy - y0)[:, None, None]; fx = (sx - x0)[None, :, None]
    y1 = np.minimum(y0 + 1, H - 1); x1 = np.minimum(x0 + 1, W - 1)
    p = img.astype(np.float32)
    top = p[y0][:, x0] * (1 - fx) + p[y0][:, x1] * fx
    bot = p[y1][:, x0] * (1 - fx) + p[y1][:, x1] * fx
    return np.floor(top * (1 - fy) + bot * fy + 0.5)


if __name__ == "__main__":
    main()
