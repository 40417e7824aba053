"""Parity of the batched landmark refinement (K13 landmark_update_kernel through the C ABI) with the CPU oracle's
restatement of Landmark::update (reference src/types/landmark.cpp:66-152): world coordinates, number of updates,
outcome and iteration count BIT-EXACT (the kernel adds the per-measurement terms in measurement order and uses the
oracle's expression order without contraction).  Stated tolerance of the path: 1e-9 relative; achieved: 0."""
import numpy as np
import pytest

from oracle import tier_a
from vslam_b200 import api, synth

pytestmark = pytest.mark.gpu


def _oracle(h, max_iterations=100, max_err2=25.0):
    n = len(h["offsets"]) - 1
    world, updates = h["world"].copy(), h["number_of_updates"].copy()
    outcome, iterations = np.zeros(n, np.uint8), np.zeros(n, np.int32)
    for i in range(n):
        ms = h["measurements"][h["offsets"][i]:h["offsets"][i + 1]]
        world[i], updates[i], outcome[i], iterations[i] = tier_a.landmark_update(
            ms, h["world_to_camera"], h["camera_to_world"], h["world"][i], h["number_of_updates"][i], max_iterations, max_err2)
    return world, updates, outcome, iterations


@pytest.mark.parametrize("n,frames,seed,outliers,behind", [(300, 40, 1, 0.05, 0.0), (500, 90, 2, 0.3, 0.0),
                                                          (64, 12, 3, 0.0, 0.2), (1, 2, 4, 0.0, 0.0)])
def test_landmark_update_matches_oracle(n, frames, seed, outliers, behind):
    h = synth.landmark_histories(n, n_frames=frames, seed=seed, outlier_fraction=outliers, behind_fraction=behind)
    opt = api.LandmarkOptimizer(n, int(h["offsets"][-1]), frames)
    got = opt.update(h["offsets"], h["measurements"], h["world_to_camera"], h["camera_to_world"], h["world"],
                     h["number_of_updates"])
    want = _oracle(h)
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
    assert np.array_equal(got[1], want[1])
    assert np.array_equal(got[0], want[0])                     # bit-exact; documented tolerance 1e-9 relative
    assert opt.launch_count == 1
    if behind == 0.0 and outliers < 0.1:
        assert (got[2] == 1).mean() > 0.5
    opt.close()


def test_iteration_cap_and_kept_state():
    h = synth.landmark_histories(50, n_frames=20, seed=8, outlier_fraction=0.0)
    opt = api.LandmarkOptimizer(50, int(h["offsets"][-1]), 20)
    start = h["world"] + 2.0
    got = opt.update(h["offsets"], h["measurements"], h["world_to_camera"], h["camera_to_world"], start,
                     h["number_of_updates"], maximum_number_of_iterations=1)
    assert np.all(got[2] == 0) and np.all(got[3] == 1)
    assert np.array_equal(got[0], start) and np.array_equal(got[1], h["number_of_updates"])
    many = np.full(50, 10000, np.uint32)                          # more updates on record than inliers: state kept
    got = opt.update(h["offsets"], h["measurements"], h["world_to_camera"], h["camera_to_world"], h["world"], many)
    assert np.all(got[2] == 3) and np.array_equal(got[0], h["world"]) and np.array_equal(got[1], many)
    opt.close()


def test_full_size_partition_invariance():
    """size-independent property at a size the oracle does not finish in seconds (20 000 landmarks, ~1 M measurements):
    landmarks are independent, so the result of a landmark does not depend on which other landmarks share the call"""
    h = synth.landmark_histories(20000, n_frames=100, seed=21, outlier_fraction=0.05)
    total = int(h["offsets"][-1])
    opt = api.LandmarkOptimizer(20000, total, 100)
    args = (h["world_to_camera"], h["camera_to_world"])
    whole = opt.update(h["offsets"], h["measurements"], *args, h["world"], h["number_of_updates"])
    for sl in (slice(0, 7000), slice(7000, 20000)):
        off = h["offsets"][sl.start:sl.stop + 1]
        part = opt.update(off - off[0], h["measurements"][off[0]:off[-1]], *args, h["world"][sl], h["number_of_updates"][sl])
        for a, b in zip(part, whole):
            assert np.array_equal(a, b[sl])
    probe = np.random.default_rng(0).choice(20000, 40, replace=False)    # and a sample against the oracle
    for i in probe:
        ms = h["measurements"][h["offsets"][i]:h["offsets"][i + 1]]
        x, nu, oc, it = tier_a.landmark_update(ms, *args, h["world"][i], h["number_of_updates"][i])
        assert np.array_equal(x, whole[0][i]) and nu == whole[1][i] and oc == whole[2][i] and it == whole[3][i]
    opt.close()


def test_argument_errors():
    h = synth.landmark_histories(4, n_frames=6, seed=2)
    opt = api.LandmarkOptimizer(4, int(h["offsets"][-1]), 6)
    args = (h["world_to_camera"], h["camera_to_world"], h["world"], h["number_of_updates"])
    bad = h["measurements"].copy()
    bad["frame"][0] = 99
    with pytest.raises(api.VslamError) as e:
        opt.update(h["offsets"], bad, *args)
    assert e.value.code == -1
    empty = h["offsets"].copy()
    empty[1] = 0
    with pytest.raises(api.VslamError):
        opt.update(empty, h["measurements"], *args)
    small = api.LandmarkOptimizer(2, 10, 6)
    with pytest.raises(api.VslamError) as e:
        small.update(h["offsets"], h["measurements"], *args)
    assert e.value.code == -3
    opt.close()
    small.close()


# ---- the device-resident landmark map (vslam_landmark_map): PoseTracker3D::_updatePoints frame by frame ------------------
def _stream_histories(h, n_frames):
    """landmark i of synth.landmark_histories was seen in the frames [n_frames - len_i, n_frames): it is born with its second
    framepoint (minimum_track_length_for_landmark_creation = 2, pose_tracker_3d.cpp:492; the constructor walks the track
    backwards, landmark.cpp:20-33) and updated by every later one"""
    births, updates = {}, {}
    for i in range(len(h["offsets"]) - 1):
        m = h["measurements"][h["offsets"][i]:h["offsets"][i + 1]]
        if len(m) < 2:
            continue
        births.setdefault(int(m["frame"][1]), []).append((i, m[[1, 0]]))
        for k in range(2, len(m)):
            updates.setdefault(int(m["frame"][k]), []).append((i, m[k]))
    return births, updates


@pytest.mark.parametrize("n,frames,seed,outliers", [(60, 12, 5, 0.05), (40, 80, 6, 0.2)])
def test_resident_landmark_map_matches_oracle_frame_by_frame(n, frames, seed, outliers):
    h = synth.landmark_histories(n, n_frames=frames, seed=seed, outlier_fraction=outliers)
    births, updates = _stream_histories(h, frames)
    w2c, c2w = h["world_to_camera"], h["camera_to_world"]
    lmap = api.LandmarkMap(n, 4 * n + 16, frames)
    ident = {}                                   # landmark of the synthetic set -> id in the map
    state = {}                                   # the oracle's (world, number_of_updates, measurements so far)
    longest = 0
    for f in range(frames):
        ups = updates.get(f, [])
        new = births.get(f, [])
        # oracle first: Landmark::update per landmark with the history grown by one measurement
        want = []
        for i, q in ups:
            world, nu, ms = state[i]
            ms = np.concatenate([ms, np.array([q], ms.dtype)])
            world, nu, outcome, its = tier_a.landmark_update(ms, w2c, c2w, world, nu)
            state[i] = (world, nu, ms)
            longest = max(longest, len(ms))
            want.append((world, nu, outcome, its))
        new_world = []
        for i, track in new:
            C = c2w[track["frame"]].reshape(-1, 3, 4)
            pts = np.einsum("fij,fj->fi", C[:, :, :3], track["camera_coordinates"]) + C[:, :, 3]
            wd = np.zeros(3)
            for p in pts:                        # `_world_coordinates += worldCoordinates()` in track order, then / size
                wd = wd + p
            wd = wd / len(pts)
            new_world.append(wd)
            state[i] = (wd, len(track), track.copy())
        offs = np.arange(0, 2 * len(new) + 1, 2, dtype=np.int32)
        tracks = np.concatenate([t for _, t in new]) if new else None
        r = lmap.update_frame(f, w2c[f], c2w[f], [ident[i] for i, _ in ups],
                              np.array([q["camera_coordinates"] for _, q in ups]).reshape(-1, 3),
                              offs if new else None, tracks, np.array(new_world) if new else None)
        for k, (world, nu, outcome, its) in enumerate(want):
            assert np.array_equal(r["world"][k], world) and r["number_of_updates"][k] == nu      # bit-exact
            assert r["outcome"][k] == outcome and r["iterations"][k] == its
        for (i, _), lid in zip(new, r["new_ids"]):
            ident[i] = int(lid)
    assert longest > 32 or frames < 40           # the 80-frame case crosses block boundaries (33+ measurements)
    ids = np.array(sorted(ident.values()), np.int32)
    assert len(lmap) == len(ids) and np.array_equal(ids, np.arange(len(ids)))
    back = {v: k for k, v in ident.items()}
    world, nu, cnt = lmap.get(ids)
    for k, lid in enumerate(ids):
        w, u, ms = state[back[int(lid)]]
        assert np.array_equal(world[k], w) and nu[k] == u and cnt[k] == len(ms)
    assert lmap.launch_count <= 2 * frames
    lmap.close()


def test_resident_landmark_map_capacity_and_arguments():
    h = synth.landmark_histories(8, n_frames=6, seed=9)
    lmap = api.LandmarkMap(4, 2, 6)              # 4 landmarks, 2 blocks in the pool
    track = h["measurements"][:2][::-1].copy()
    I = np.hstack([np.eye(3), np.zeros((3, 1))])
    with pytest.raises(api.VslamError):          # unknown id
        lmap.update_frame(0, I, I, [0], np.ones((1, 3)))
    offs = np.array([0, 2, 4, 6], np.int32)
    with pytest.raises(api.VslamError):          # three landmarks need three blocks
        lmap.update_frame(0, I, I, [], [], offs, np.concatenate([track] * 3), np.ones((3, 3)))
    with pytest.raises(api.VslamError):          # capacity of the map itself
        lmap.update_frame(0, I, I, [], [], np.arange(0, 11, 2, dtype=np.int32), np.concatenate([track] * 5), np.ones((5, 3)))
    lmap.close()
