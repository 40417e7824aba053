#!/usr/bin/env python3
"""Per-source-line instruction histogram of one kernel of an .ncu-rep captured with --import-source on.
Usage: tools/ncu_lines.py <file.ncu-rep> [kernel-regex] [min-share-percent]"""
import csv
import subprocess
import sys


def num(s):
    try:
        return int(s.replace(",", ""))
    except ValueError:
        return 0


def main():
    path = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 else None
    floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
    cmd = ["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if kern:
        cmd += ["--kernel-name", "regex:" + kern]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hi]
    ie, te, ss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index(
        "Warp Stall Sampling (All Samples)")
    lines, tot, samp, cur_file = [], 0, 0, ""
    for r in rows[hi + 1:]:
        if r and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        if r and r[0].isdigit() and len(r) > te:
            n = num(r[ie])
            tot += n
            samp += num(r[ss])
            lines.append((cur_file, int(r[0]), n, num(r[te]), num(r[ss]), r[1]))
    print("total warp instructions %d, samples %d" % (tot, samp))
    for f, ln, n, t, s, src in lines:
        if n > tot * floor / 100 or s > samp * floor / 100:
            print("%4d %6.2f%% inst  %5.1f thr/warp  %6.2f%% samples  %s" % (ln, 100 * n / max(tot, 1), t / max(n, 1),
                                                                            100 * s / max(samp, 1), src.strip()[:100]))


if __name__ == "__main__":
    main()
