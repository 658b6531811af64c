"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the EfficientNMS stage + parse_output.

The reference runs TensorRT's EfficientNMS_TRT plugin inside its engine
(/root/reference/src/yolo_engine.cpp:33, README.md:25) and reads back num_dets / det_boxes /
det_scores / det_classes (:53-57, 63-69), then scales boxes to the source frame in
parse_output (:202-220).  The plugin source (TensorRT OSS plugin/efficientNMSPlugin) is a
third-party dependency that is NOT vendored in /root/reference and its parameters live in the
absent ONNX, so this file restates the published algorithm with the YOLOv7-end2end export
defaults (score_threshold 0.25, iou_threshold 0.45, max_output_boxes 100) and fixes the total
order the plugin leaves unspecified:

    candidates = every (anchor a, class c) with score > score_thr          (strict)
    order      = score descending, then flat index a*nc+c ascending
    pre-NMS    = only the first `max_candidates` (4096) in that order enter NMS
    suppress   = candidate j is dropped by an already-kept i iff class_i == class_j and
                 IoU(box_i, box_j) > iou_thr                               (strict)
    stop       = after `max_det` kept

IoU is computed in FP32 with individually rounded operations, the arithmetic the CUDA kernel
repeats, so kept indices are bit-exact given identical boxes and scores.
Parity status: "parity unpinned" by the reference (no NMS test); cross-checked against
cv2.dnn.NMSBoxesBatched in tests/test_oracle_cpu.py.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

SCORE_THR = 0.25
IOU_THR = 0.45
MAX_DET = 100
MAX_CAND = 4096
N_CLASSES = 14
UNKNOWN = 14   # ArmorClass::UNKNOWN, /root/reference/include/irmv_detection/armor.hpp:7


def iou_f32(a: np.ndarray, b: np.ndarray) -> np.float32:
    f = np.float32
    ix1 = max(f(a[0]), f(b[0]))
    iy1 = max(f(a[1]), f(b[1]))
    ix2 = min(f(a[2]), f(b[2]))
    iy2 = min(f(a[3]), f(b[3]))
    iw = max(f(ix2 - ix1), f(0))
    ih = max(f(iy2 - iy1), f(0))
    inter = f(iw * ih)
    area_a = f(f(a[2] - a[0]) * f(a[3] - a[1]))
    area_b = f(f(b[2] - b[0]) * f(b[3] - b[1]))
    union = f(f(area_a + area_b) - inter)
    if not union > f(0):
        return f(0)
    return f(inter / union)


def nms(boxes: np.ndarray, scores: np.ndarray, score_thr=SCORE_THR, iou_thr=IOU_THR,
        max_det=MAX_DET, max_cand=MAX_CAND):
    """boxes f32[A,4] xyxy, scores f32[A,nc] -> (flat_idx i32[n], boxes f32[n,4], scores f32[n], cls i32[n])."""
    boxes = np.ascontiguousarray(boxes, np.float32)
    scores = np.ascontiguousarray(scores, np.float32)
    A, nc = scores.shape
    flat = scores.reshape(-1)
    cand = np.nonzero(flat > np.float32(score_thr))[0]
    # score desc, flat index asc (stable sort on an index-ascending list)
    order = cand[np.argsort(-flat[cand], kind="stable")]
    order = order[:max_cand]
    keep = []
    iou_thr = np.float32(iou_thr)
    for j in order:
        aj, cj = divmod(int(j), nc)
        ok = True
        for i in keep:
            ai, ci = divmod(int(i), nc)
            if ci == cj and iou_f32(boxes[ai], boxes[aj]) > iou_thr:
                ok = False
                break
        if ok:
            keep.append(int(j))
            if len(keep) >= max_det:
                break
    keep = np.asarray(keep, np.int32)
    a = keep // nc
    c = keep % nc
    return keep, boxes[a].reshape(-1, 4), flat[keep].astype(np.float32), c.astype(np.int32)


def parse_output(det_boxes: np.ndarray, det_scores: np.ndarray, det_classes: np.ndarray,
                 src_w: int, src_h: int, nc: int = N_CLASSES):
    """/root/reference/src/yolo_engine.cpp:202-220: scale to the source frame, enum_cast class."""
    sx = np.float32(src_w) / np.float32(640)
    sy = np.float32(src_h) / np.float32(640)
    out = det_boxes.astype(np.float32).copy().reshape(-1, 4)
    out[:, 0] *= sx
    out[:, 1] *= sy
    out[:, 2] *= sx
    out[:, 3] *= sy
    cls = np.where((det_classes >= 0) & (det_classes < nc), det_classes, UNKNOWN).astype(np.int32)
    return out, det_scores.astype(np.float32), cls
