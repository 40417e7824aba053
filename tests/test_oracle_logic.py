"""Reference-owned logic of the oracle: detector regions, threshold controller, triangulation threshold,
stereo scan semantics (cursor, disparity, ties, epipolar passes), bin rule -- on hand-made cases."""
import numpy as np

from oracle import tier_a

CAM = tier_a.StereoCamera(718.856, 718.856, 607.1928, 185.2157, -386.1448)


def test_detector_regions_euroc_2x2():
    r = tier_a.detector_regions(480, 752, 2, 2)
    got = [(int(q["x"]), int(q["y"]), int(q["w"]), int(q["h"])) for q in r]
    assert got == [(0, 0, 378, 242), (374, 0, 378, 242), (0, 238, 378, 242), (374, 238, 378, 242)]


def test_detector_regions_single_and_3x3():
    r = tier_a.detector_regions(376, 1241, 1, 1)
    assert (int(r[0]["x"]), int(r[0]["y"]), int(r[0]["w"]), int(r[0]["h"])) == (0, 0, 1241, 376)
    r = tier_a.detector_regions(376, 1241, 3, 3)
    # central regions get the doubled overlap (base_framepoint_generator.cpp:279-281,287-289)
    assert (int(r[4]["x"]), int(r[4]["y"]), int(r[4]["w"]), int(r[4]["h"])) == (412, 123, 417, 129)
    for q in r:
        assert q["x"] >= 0 and q["y"] >= 0 and q["x"] + q["w"] <= 1241 and q["y"] + q["h"] <= 376


def test_bin_grid():
    assert tier_a.bin_grid(376, 1241, 15) == (26, 83)
    assert tier_a.bin_grid(376, 1241, 25) == (16, 50)
    assert tier_a.bin_grid(480, 752, 20) == (25, 38)
    assert tier_a.bin_grid(1080, 1920, 23) == (47, 84)


def test_threshold_controller():
    f = tier_a.lib().orc_threshold_proposal
    # too few points: lower by max(delta,-maxchg)*thr but at least 1, clamp at minimum
    assert f(40.0, 1000, 2158.0, 0.1, 0.1, 20.0, 100.0) == 36.0
    assert f(20.0, 1000, 2158.0, 0.1, 0.1, 20.0, 100.0) == 20.0
    assert f(5.0, 2000, 2158.0, 0.01, 0.1, 1.0, 100.0) == 4.0        # change*thr = -0.37 -> -1
    # too many points: raise, at least 1, clamp at maximum
    assert f(20.0, 3230, 2158.0, 0.1, 0.1, 20.0, 100.0) == 22.0
    assert f(99.5, 9999, 2158.0, 0.1, 0.1, 20.0, 100.0) == 100.0
    assert f(5.0, 2400, 2158.0, 0.1, 0.5, 1.0, 100.0) == 6.0         # 0.112*5 < 1 -> +1
    # inside tolerance: unchanged
    assert f(33.0, 2200, 2158.0, 0.1, 0.1, 20.0, 100.0) == 33.0
    # L and R proposals averaged then rint (half to even): (6 + 7)/2 = 6.5 -> 6 ; (5 + 6)/2 = 5.5 -> 6
    thr = tier_a.adjust_thresholds([6.0, 5.0], [2158, 2158], [2400, 2400], 2158, 0.1, 0.1, 1, 100)
    assert list(thr) == [6.0, 6.0]
    thr = tier_a.adjust_thresholds([21.0], [2158], [2500], 2158, 0.1, 0.1, 20, 100)   # (21+23.1)/2=22.05
    assert thr[0] == 22.0


def test_triangulation_threshold():
    assert tier_a.triangulation_threshold(True, 100, 2158, 51.2) == 25.6
    assert tier_a.triangulation_threshold(True, 100, 2158, 20.0) == 20.0
    assert tier_a.triangulation_threshold(False, 5000, 2158, 51.2) == 51.2
    assert tier_a.triangulation_threshold(False, 100, 2158, 51.2) == 25.6
    assert tier_a.triangulation_threshold(False, 1500, 2158, 60.0) == (1500 / 2158) * 60.0


def _feat(rows_cols_desc):
    f = np.zeros(len(rows_cols_desc), tier_a.FEATURE)
    for i, (r, c, d) in enumerate(rows_cols_desc):
        f[i]["x"], f[i]["y"], f[i]["row"], f[i]["col"], f[i]["index"] = c, r, r, c, i
        f[i]["desc"] = d
    return f


def _desc(nbits):
    d = np.zeros(32, np.uint8)
    for b in range(nbits):
        d[b // 8] |= 1 << (b % 8)
    return d


def _run(fl, fr, thr=25.6, min_disp=1.0, off=0, binning=False, bin_size=15, tracked=None):
    return tier_a.stereo_compute(fl, fr, CAM, thr, min_disp, off, binning, bin_size, 376, 1241, tracked)


def test_stereo_scan_basic_and_tie_lowest_column_wins():
    fl = _feat([(50, 200, _desc(0))])
    fr = _feat([(50, 150, _desc(3)), (50, 160, _desc(3)), (50, 170, _desc(5)), (50, 210, _desc(0))])
    r = _run(fl, fr)
    m = r["matches"]
    assert len(m) == 1 and m[0]["index_right"] == 0 and m[0]["distance"] == 3   # strict '<': first min wins
    assert m[0]["epipolar_offset"] == 0
    assert len(r["remaining_left"]) == 0 and len(r["remaining_right"]) == 3


def test_stereo_scan_stops_at_negative_disparity_and_threshold_is_strict():
    fl = _feat([(50, 200, _desc(0))])
    fr = _feat([(50, 201, _desc(0))])               # col_R > col_L: never evaluated
    assert len(_run(fl, fr)["matches"]) == 0
    fr = _feat([(50, 100, _desc(26))])              # 26 !< 25.6
    assert len(_run(fl, fr)["matches"]) == 0
    fr = _feat([(50, 100, _desc(25))])              # 25 < 25.6
    assert len(_run(fl, fr)["matches"]) == 1
    fr = _feat([(50, 100, _desc(25))])
    assert len(_run(fl, fr, thr=25.0)["matches"]) == 0   # 25 !< 25.0


def test_stereo_scan_monotone_cursor_blocks_earlier_right_features():
    # L0 takes R1 (better than R0); cursor moves past R1, so L1 can no longer see R0.
    fl = _feat([(50, 300, _desc(0)), (50, 320, _desc(8))])
    fr = _feat([(50, 100, _desc(8)), (50, 120, _desc(1)), (50, 310, _desc(9))])
    r = _run(fl, fr)
    m = r["matches"]
    assert [(int(a["index_left"]), int(a["index_right"]), int(a["distance"])) for a in m] == [(0, 1, 1), (1, 2, 1)]


def test_stereo_scan_minimum_disparity_does_not_consume():
    # best candidate for L0 has disparity 0 (< 1): skipped WITHOUT moving the cursor (cpp:358-361),
    # so L1 still sees the same right feature.
    fl = _feat([(50, 300, _desc(0)), (50, 330, _desc(0))])
    fr = _feat([(50, 300, _desc(0))])
    r = _run(fl, fr)
    m = r["matches"]
    assert len(m) == 1 and m[0]["index_left"] == 1 and m[0]["index_right"] == 0


def test_stereo_rows_are_independent_and_unsorted_input_rows_skip():
    fl = _feat([(10, 100, _desc(0)), (20, 100, _desc(0)), (30, 100, _desc(0))])
    fr = _feat([(5, 50, _desc(0)), (20, 50, _desc(0)), (25, 60, _desc(0)), (30, 90, _desc(1))])
    m = _run(fl, fr)["matches"]
    assert [(int(a["index_left"]), int(a["index_right"])) for a in m] == [(1, 1), (2, 3)]


def test_epipolar_offsets_pass_order_and_pruning():
    # offsets visited 0, +1, -1 (cpp:45-50); a left feature at row r matches right rows r, r-1, r+1 in that order
    fl = _feat([(50, 300, _desc(0)), (60, 300, _desc(0)), (70, 300, _desc(0))])
    fr = _feat([(49, 200, _desc(1)), (60, 200, _desc(2)), (71, 200, _desc(3))])
    m = _run(fl, fr, off=1)["matches"]
    got = [(int(a["index_left"]), int(a["index_right"]), int(a["epipolar_offset"])) for a in m]
    assert got == [(1, 1, 0), (0, 0, 1), (2, 2, -1)]
    # a feature matched in pass 0 is pruned and cannot be re-used in pass +1
    fl = _feat([(50, 300, _desc(0)), (51, 300, _desc(0))])
    fr = _feat([(50, 200, _desc(0))])
    m = _run(fl, fr, off=1)["matches"]
    assert [(int(a["index_left"]), int(a["index_right"])) for a in m] == [(0, 0)]


def test_triangulation_formula():
    out = tier_a.triangulate(CAM, np.float32(700.0), np.float32(100.0), np.float32(680.0), np.float32(100.0))
    z = -386.1448 / (680.0 - 700.0)
    assert out[2] == z
    assert out[0] == ((1 / 718.856) * (700.0 - 607.1928)) * z
    assert out[1] == ((1 / 718.856) * ((100.0 + 100.0) / 2.0 - 185.2157)) * z


def test_bin_rule_is_order_dependent_partial_order():
    # same bin (rint(row/15), rint(col/15)) = (3, 20); candidates in emission order (row, col)
    # A: disp 10, dist 5 ; B: disp 20, dist 5 -> replaces A ; C: disp 30, dist 6 -> does NOT replace (dist worse)
    d0 = _desc(0)
    fl = _feat([(44, 296, d0), (45, 300, d0), (46, 304, d0)])
    fr = _feat([(44, 286, _desc(5)), (45, 280, _desc(5)), (46, 274, _desc(6))])
    r = _run(fl, fr, binning=True)
    assert len(r["matches"]) == 3
    assert list(r["winners"]) == [1]
    # reversed quality order: first point is already the best, nothing replaces it
    fr = _feat([(44, 266, _desc(5)), (45, 280, _desc(5)), (46, 294, _desc(4))])
    r = _run(fl, fr, binning=True)
    assert list(r["winners"]) == [0]


def test_bin_preload_blocks_and_winner_order_is_row_major_bins():
    d0 = _desc(0)
    fl = _feat([(45, 300, d0), (45, 900, d0), (200, 100, d0)])
    fr = _feat([(45, 280, d0), (45, 880, d0), (200, 80, d0)])
    tracked = np.zeros(1, tier_a.TRACKED)
    tracked[0]["row"], tracked[0]["col"], tracked[0]["has_previous"] = 44, 898, 1     # same bin as L1
    r = _run(fl, fr, binning=True, tracked=tracked)
    assert len(r["matches"]) == 3
    assert list(r["winners"]) == [0, 2]            # L1 blocked by the tracked point; order = bin row-major
    r = _run(fl, fr, binning=False)
    assert list(r["winners"]) == [0, 1, 2]
