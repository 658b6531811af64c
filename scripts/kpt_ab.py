import os, sys, statistics, torch
sys.path.insert(0, "/root/repo")
import irmv_detection_b200 as irmv
from irmv_detection_b200 import synth, weights
path = "/tmp/sh.irmw"
weights.write_random(path, 0, arch="shufflenetv2-pose")
fr = torch.from_numpy(synth.frames_from_base(synth.load_base(), 64, seed=2)).cuda()
for fuse in (True, False):
    eng = irmv.YoloEngine(path, (1280, 1024), max_batch=64, sub_batch=64, num_lanes=1, fuse_units=fuse)
    for _ in range(5):
        eng.enqueue_batch_device(fr.data_ptr(), 64); eng.sync()
    ms = []
    for i in range(20):
        eng.enqueue_batch_device(fr.data_ptr(), 64); ms.append(eng.sync())
    k, st = eng.profile_stages(fr.data_ptr(), 64)
    print("fuse", fuse, "ms/64", statistics.median(ms), st, eng.kernel_launches(64), flush=True)
    for o in eng.profile_ops(64) if hasattr(eng, "profile_ops") else []:
        pass
    eng.close()
