"""per-round and fixed cost of the fused converge() at the problem size of a tracked frame"""
import dataclasses, sys, time
sys.path.insert(0, ".")
import numpy as np
from vslam_b200 import api, configs, synth
cam = synth.camera("kitti")
for n in (720, 2000, 100000):
    c = synth.correspondences(n, "stereouv", cam, seed=3)
    T0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    res = []
    for cap in (1, 2, 4, 8, 1000):
        acfg = dataclasses.replace(configs.KITTI_ALIGNER, maximum_number_of_iterations=cap)
        al = api.StereoUVAligner(acfg, max_points=max(n, 4096))
        al.initialize(c["moving"], c["fixed"], c["omega"], c["wt"], cam.K, cam.baseline, cam.rows, cam.cols, T0)
        al.converge(fused=True)
        t0 = time.perf_counter()
        reps = 200 if n < 10000 else 50
        for _ in range(reps):
            al.setPreviousToCurrent(T0)
            al.converge(fused=True)
        us = (time.perf_counter() - t0) / reps * 1e6
        res.append((cap, al.number_of_rounds, us))
        al.close()
    print("n=%d:" % n, "  ".join("cap %d -> %d rounds %.1f us" % r for r in res))
    (c1, r1, t1), (c2, r2, t2) = res[0], res[3]
    print("   per round %.2f us, fixed %.1f us" % ((t2 - t1) / (r2 - r1), t1 - r1 * (t2 - t1) / (r2 - r1)))
