import sys, time
sys.path.insert(0, ".")
import numpy as np
from vslam_b200 import api, synth
n, frames = 1000, 60
h = synth.landmark_histories(n, n_frames=frames, seed=5, outlier_fraction=0.05)
lmap = api.LandmarkMap(n, 8 * n, frames + 64)
for f in range(frames):
    lmap.set_frame_pose(f, h["world_to_camera"][f], h["camera_to_world"][f])
born = [h["measurements"][h["offsets"][i]:h["offsets"][i + 1] - 1][::-1] for i in range(n)]
born = [bm if len(bm) else h["measurements"][h["offsets"][i]:h["offsets"][i + 1]] for i, bm in enumerate(born)]
offs = np.concatenate([[0], np.cumsum([len(bm) for bm in born])]).astype(np.int32)
last = frames - 1
W, Cw = h["world_to_camera"][last], h["camera_to_world"][last]
r0 = lmap.update_frame(last, W, Cw, [], [], offs, np.concatenate(born), h["world"])
ids = r0["new_ids"]
cam = np.array([h["measurements"][h["offsets"][i + 1] - 1]["camera_coordinates"] for i in range(n)])
def t(fn, reps=50):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e6
print("empty call us", t(lambda: lmap.update_frame(last, W, Cw, [], [])))
print("10 landmarks us", t(lambda: lmap.update_frame(last, W, Cw, ids[:10], cam[:10])))
print("100 landmarks us", t(lambda: lmap.update_frame(last, W, Cw, ids[:100], cam[:100])))
print("1000 landmarks us", t(lambda: lmap.update_frame(last, W, Cw, ids, cam), 20))
r = lmap.update_frame(last, W, Cw, ids, cam)
print("iterations: mean %.2f max %d, landmarks at the cap: %d" % (r["iterations"].mean(), r["iterations"].max(), (r["iterations"] >= 100).sum()))
print("1000 landmarks, 1 iteration cap us", t(lambda: lmap.update_frame(last, W, Cw, ids, cam, maximum_number_of_iterations=1), 10))
