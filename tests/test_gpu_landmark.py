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
