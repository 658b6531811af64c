"""One eager replay of the pipeline for ncu: `python scripts/profile_replay.py [frames] [warmups]`.
Prints the number of kernel launches per replay so ncu's -s/-c can be set."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import irmv_detection_b200 as irmv
    from irmv_detection_b200 import weights
    import bench
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    pose = "pose" in sys.argv[3:]                            # keypoint variant (72-conv weight file)
    w = "/tmp/profile_seed0_pose.irmw" if pose else "/tmp/profile_seed0.irmw"
    weights.write_random(w, 0, pose=pose)
    dev = torch.device("cuda", 0)
    frames = bench.make_bayer_frames_device(n, 0, dev)
    torch.cuda.synchronize()
    eng = irmv.YoloEngine(w, (1280, 1024), chan_order=irmv.CH_BAYER_RGGB, max_batch=n, sub_batch=n, num_lanes=1,
                          use_graph=False)
    if "armors" in sys.argv[3:]:                            # light-bar extraction between NMS and PnP
        eng.enable_armors()
    eng.enable_pnp(bench.K_CAM, bench.D_CAM, (0.5, 480 / 1024))
    for _ in range(warm + 1):
        eng.enqueue_batch_device(frames.data_ptr(), n)
        ms = eng.sync()
    print("launches per replay", eng.kernel_launches(n), "device ms", ms, "frames/s", n / (ms * 1e-3), "keypoints", eng.has_keypoints())
    import json
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(eng.describe_ops(), open(os.path.join(ROOT, "gpurun_out", "ops.json"), "w"))
    k, st = eng.profile_stages(frames.data_ptr(), n)
    print("stages", k, st)
    eng.close()


if __name__ == "__main__":
    main()
