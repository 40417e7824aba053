"""Launches every kernel of libvslam_b200.so once or twice through the C ABI, standalone (no oracle, no pytest).

Used as the target of
  compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tools/kernel_zoo.py --small
  ncu --set full ... python tools/kernel_zoo.py --profile
(`--small`: the smallest case of every kernel, so the sanitizers' 10-100x slowdown stays within a minute;
 `--profile`: the sizes the roofline numbers are quoted on -- 4 M correspondences for linearize_kernel<0|1>,
 a 64-pair KITTI batch for the framepoint kernels, 20 000 landmarks for landmark_update_kernel.)
Every stage prints the counts it produced: a sanitizer run that changed a result would show up as a different line.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vslam_b200 import api, configs, synth  # noqa: E402


def previous_points_from(fps, desc_left, desc_right, landmark_every=2):
    """vslam_previous_point records from the framepoints compute() returned (what track() reads of the previous frame)"""
    p = np.zeros(len(fps), api.PREVIOUS_POINT)
    p["camera_left"] = fps["camera"]
    p["world"] = fps["camera"]
    p["descriptor_left"] = desc_left[fps["index_left"]]
    p["descriptor_right"] = desc_right[fps["index_right"]]
    p["epipolar_offset"] = fps["epipolar_offset"]
    p["has_landmark"][::landmark_every] = 1
    p["keypoint_size"] = 7.0
    return p


def fpg_sequence(cfg_name, frames, seed=5):
    cfg = configs.BY_NAME[cfg_name]
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, seed, max_frames=frames + 1)
    gen = api.StereoFramePointGenerator(cfg, cam)
    prev = None
    Tk = np.hstack([np.eye(3), np.zeros((3, 1))])
    Tk[0, 3] = -(-cam.bx / cam.fx) / 4          # previous -> current of BandWorld: a camera x-translation of B/4 per frame
    for k in range(frames):
        left, right = world.pair(k)
        nl, nr = gen.initialize(left, right, k == 0)
        kl, dl = gen.features(0)
        kr, dr = gen.features(1)
        line = "%s frame %d: features %d/%d" % (cfg_name, k, nl, nr)
        if prev is not None and len(prev):
            got = gen.track(prev, Tk, k == 1, 50 if k == 1 else 20, 38.4)
            line += ", tracks %d lost %d" % (len(got["tracks"]), len(got["lost"]))
            lost = prev[got["lost"]]
            if len(lost):
                lost["world"] = lost["camera_left"]      # world == previous camera frame here: world -> camera = Tk
                rec = gen.recover_points(lost, Tk, 51.2)
                line += ", recovered %d" % len(rec)
            fps = gen.compute(api.TRACKED_FROM_LAST_TRACK)
        else:
            fps = gen.compute()
        line += ", matches %d, framepoints %d" % (gen.number_of_matches, len(fps))
        print(line, flush=True)
        new = fps[fps["index_left"] >= 0]
        prev = previous_points_from(new, dl, dr)
    gen.close()


def frame_steps(cfg_name, frames, seed=7):
    """vslam_fpg_frame_step: the tracked frame as one graph (compact_frame / track_search / track_resolve / track_emit /
    converge_cluster with initialize + prune / match / select_strips<32> / frame_assemble)"""
    cfg, acfg = configs.BY_NAME[cfg_name], configs.ALIGNER_BY_NAME[cfg_name]
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, seed, max_frames=frames + 1)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.frame_step_reset()
    T = np.hstack([np.eye(3), np.zeros((3, 1))])
    T[0, 3] = -(-cam.bx / cam.fx) / 4
    for k in range(frames):
        left, right = world.pair(k)
        r = gen.frame_step(left, right, k == 0, T, acfg, False, 15, 25.6, publish_frame_points=True)
        print("%s fused frame %d: previous %d tracks %d new %d rounds %d" % (cfg_name, k, r["n_previous"], r["n_tracks"],
                                                                           r["n_new_points"], r["aligner_rounds"]), flush=True)
    gen.close()


def fpg_batch(cfg_name, n, rounds=2):
    cfg = configs.BY_NAME[cfg_name]
    cam = synth.camera(cfg.camera)
    left, right = synth.band_world_batch(cfg.camera, range(100, 100 + min(n, 8)))
    if n > len(left):
        reps = (n + len(left) - 1) // len(left)
        left, right = np.tile(left, (reps, 1, 1))[:n], np.tile(right, (reps, 1, 1))[:n]
    gen = api.StereoFramePointGenerator(cfg, cam, max_batch=n)
    out, counts = gen.batch_process(left, right, True)
    T = np.hstack([np.eye(3), np.array([[0.0], [0.0], [0.15]])])
    gen.batch_linearize(n, T, configs.KITTI_FAST_ALIGNER, ignore_outliers=False, rounds=rounds)
    systems = gen.batch_systems(n)
    print("%s batch of %d: framepoints %d, inliers %d" % (cfg_name, n, int(counts.sum()),
                                                          int(sum(s["inliers"] for s in systems))), flush=True)
    gen.close()


def aligner(kind, n):
    cam = synth.camera("kitti")
    c = synth.correspondences(n, kind, cam)
    cls = api.StereoUVAligner if kind == "stereouv" else api.UVDAligner
    al = cls(configs.KITTI_FAST_ALIGNER, max_points=n)
    omega = c["omega"] if kind == "stereouv" else np.stack([c["omega_uv"], c["omega_d"]], 1)
    al.initialize(c["moving"], c["fixed"], omega, c["wt"], cam.K, cam.baseline, cam.rows, cam.cols)
    s = al.linearize(False)
    s2 = al.linearize(True)
    al.oneRound(False)
    al.setPreviousToCurrent(np.hstack([np.eye(3), np.zeros((3, 1))]))
    r = al.converge(fused=True)
    print("%s aligner n=%d: inliers %d / %d, converged %s after %d rounds, inliers %d" % (
        kind, n, s["inliers"], s2["inliers"], al.has_system_converged, al.number_of_rounds, r["inliers"]), flush=True)
    al.close()


def landmarks(n, frames):
    h = synth.landmark_histories(n, n_frames=frames, seed=4, outlier_fraction=0.05)
    opt = api.LandmarkOptimizer(n, int(h["offsets"][-1]), frames)
    got = opt.update(h["offsets"], h["measurements"], h["world_to_camera"], h["camera_to_world"], h["world"],
                     h["number_of_updates"])
    print("landmarks n=%d: adopted %d" % (n, int((got[2] == 1).sum())), flush=True)
    opt.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    if a.profile:
        fpg_batch("kitti_fast", 64, rounds=2)
        fpg_sequence("kitti", 3)
        frame_steps("kitti", 4)
        aligner("stereouv", 4_000_000)
        aligner("uvd", 4_000_000)
        landmarks(20000, 100)
    else:
        fpg_sequence("euroc", 3)
        frame_steps("euroc", 3)
        fpg_batch("kitti_fast", 2)
        aligner("stereouv", 3000)
        aligner("uvd", 3000)
        landmarks(40, 8)
    print("kernel zoo done", flush=True)


if __name__ == "__main__":
    main()
