import sys, time
sys.path.insert(0, '.')
import numpy as np
from vslam_b200 import api, configs, synth
for name in ("kitti", "euroc"):
    cfg = configs.BY_NAME[name]; cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 7, max_frames=40)
    frames = []
    for k in range(24):
        l, r = world.pair(k); pl, pr = api.pinned_empty(l.shape), api.pinned_empty(r.shape); pl[:], pr[:] = l, r; frames.append((pl, pr))
    gen = api.StereoFramePointGenerator(cfg, cam)
    for k in range(4):
        gen.initialize(frames[k][0], frames[k][1], k == 0); gen.compute()
    gen.set_profiling(-1); gen.set_profiling(1)
    t_init = t_comp = 0
    for k in range(4, 24):
        t0 = time.perf_counter(); gen.initialize(frames[k][0], frames[k][1], False); t1 = time.perf_counter(); gen.compute(); t2 = time.perf_counter()
        t_init += t1 - t0; t_comp += t2 - t1
    prof = gen.kernel_profile()
    print(name, 'initialize ms', t_init/20*1e3, 'compute ms', t_comp/20*1e3)
    print({k: round(v[0]/20*1e3, 1) for k, v in prof.items()}, 'us per frame')
    gen.close()
