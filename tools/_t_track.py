import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from vslam_b200 import api, configs, synth
for name in ("kitti", "euroc"):
    cfg = configs.BY_NAME[name]; cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 7, max_frames=40)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.set_profiling(1)
    T = np.hstack([np.eye(3), np.zeros((3, 1))]); T[0, 3] = -(-cam.bx / cam.fx) / 4
    prev = None
    tt = []
    for k in range(12):
        l, r = world.pair(k)
        gen.initialize(l, r, k == 0)
        parts = []
        if prev is not None:
            t0 = time.perf_counter(); res = gen.track(prev, T, False, 15, 25.6); t1 = time.perf_counter()
            tt.append((t1 - t0) * 1e3)
            parts.append(res["tracks"]); new = gen.compute(api.TRACKED_FROM_LAST_TRACK)
        else:
            new = gen.compute()
        parts.append(new)
        _, dl = gen.features(0); _, dr = gen.features(1)
        prev = api.make_previous_points(parts, dl, dr)
    prof = gen.kernel_profile()
    print({k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in prof.items()}, "us per launch")
    print(name, "track() ms per call:", np.round(tt, 3), "prev", len(prev))
