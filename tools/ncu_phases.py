#!/usr/bin/env python3
"""Executed warp instructions of fast_nms_kernel per PHASE of the kernel (pre-test, candidate list, arc test, NMS,
publish, per-tile fixed cost), from the per-source-line counters of an .ncu-rep captured with --import-source on.
The phases are found from the `// ---- phase N` markers and the helper functions of csrc/fast.cu, so the split follows
the source it was captured from.  Usage: tools/ncu_phases.py <file.ncu-rep> <n_pixels_per_launch> [fast.cu]"""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(s):
    try:
        return int(s.replace(",", ""))
    except ValueError:
        return 0


def phase_ranges(path):
    src = open(path).read().splitlines()

    def find(pattern, start=0):
        for i in range(start, len(src)):
            if re.search(pattern, src[i]):
                return i + 1
        raise SystemExit("marker not found: " + pattern)
    k0 = find(r"__global__ void __launch_bounds__\(256\) fast_nms_kernel")
    marks = [("0 per-tile fixed cost (index math, TMA issue, clears, barrier wait)", k0),
             ("1 compass pre-test (all pixels, 4 per thread)", find(r"// ---- phase 1", k0)),
             ("2 warp scan + candidate list", find(r"ONE warp scan places", k0)),
             ("3 arc test + corner score (candidates)", find(r"// ---- phase 2", k0)),
             ("4 non-maximum suppression", find(r"// ---- phase 3", k0)),
             ("5 publish mask words + raw count", find(r"// ---- phase 4", k0))]
    end = find(r"^// K2", k0)
    ranges = []
    for (name, lo), nxt in zip(marks, [m[1] for m in marks[1:]] + [end]):
        ranges.append((name, lo, nxt - 1))
    helpers = [("1 compass pre-test (all pixels, 4 per thread)", find(r"uint32_t absdiff_gt\("), find(r"uint32_t smem_u32\(") - 1),
               ("3 arc test + corner score (candidates)", find(r"bool arc9\("), find(r"uint32_t absdiff_gt\(") - 1)]
    return ranges + helpers


def main():
    rep, pixels = sys.argv[1], float(sys.argv[2])
    src = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "vslam-pose-estimation-framework_b200", "csrc", "fast.cu")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:fast_nms_kernel"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hi]
    ie, ss = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    per, other, in_main = {}, [0, 0], True
    for r in rows[hi + 1:]:
        if r and r[0] == "File Path":
            in_main = r[1].endswith("fast.cu")
        if r and r[0].isdigit() and len(r) > ie:
            if in_main:
                a, b = per.get(int(r[0]), (0, 0))
                per[int(r[0])] = (a + num(r[ie]), b + num(r[ss]))
            else:
                other[0] += num(r[ie])
                other[1] += num(r[ss])
    ranges = phase_ranges(src)
    acc = {}
    for line, (a, b) in per.items():
        name = next((n for n, lo, hi2 in ranges if lo <= line <= hi2), "6 other lines of fast.cu")
        x, y = acc.get(name, (0, 0))
        acc[name] = (x + a, y + b)
    acc["7 inlined CUDA headers (atomics, shuffles)"] = tuple(other)
    tot = sum(a for a, _ in acc.values())
    st = max(1, sum(b for _, b in acc.values()))
    print("fast_nms_kernel: %d executed warp instructions for %.0f pixels = %.3f warp instructions per pixel" % (tot, pixels, tot / pixels))
    print("%-72s %8s %10s %9s" % ("phase", "inst %", "inst/pixel", "samples %"))
    for name in sorted(acc):
        a, b = acc[name]
        print("%-72s %7.2f%% %10.3f %8.2f%%" % (name, 100.0 * a / tot, a / pixels, 100.0 * b / st))


if __name__ == "__main__":
    main()
