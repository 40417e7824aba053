#!/usr/bin/env python3
"""Extracts the BRIEF-32 test table from opencv_contrib's modules/xfeatures2d/src/generated_32.i (not shipped here).

That file defines `#define SMOOTHED(y,x) smoothedSum(sum, pt, y, x, ...)` and 32 statements of the form
    desc[i] = (uchar)(((SMOOTHED(y0, x0) < SMOOTHED(y1, x1)) << 7) + ((SMOOTHED(..) < SMOOTHED(..)) << 6) + ... );
This tool returns / writes the 256 x 4 int8 table (y0, x0, y1, x1), test 8 i + k = the term shifted by (7 - k) of
desc[i]: the layout vslam_fpg_config.brief_tests expects.

usage: parse_brief_generated.py generated_32.i brief_tests.i8      (raw 1024 bytes)
"""
import re
import sys

import numpy as np

_TERM = re.compile(r"SMOOTHED\(\s*(-?\d+)\s*,\s*(-?\d+)\s*\)\s*<\s*SMOOTHED\(\s*(-?\d+)\s*,\s*(-?\d+)\s*\)\s*\)\s*<<\s*(\d)")
_STMT = re.compile(r"desc\[(\d+)\]\s*=\s*\(uchar\)\((.*?)\);", re.S)


def parse(text: str) -> np.ndarray:
    table = np.zeros((256, 4), np.int8)
    seen = np.zeros(256, bool)
    for byte, body in _STMT.findall(text):
        for y0, x0, y1, x1, shift in _TERM.findall(body):
            k = 8 * int(byte) + (7 - int(shift))
            table[k] = (int(y0), int(x0), int(y1), int(x1))
            seen[k] = True
    if not seen.all():
        raise ValueError("generated_32.i: found %d of 256 tests" % seen.sum())
    return table


def render(table: np.ndarray) -> str:
    """the inverse of parse(): text in the format of generated_32.i (used by the tests)"""
    out = ["#define SMOOTHED(y,x) smoothedSum(sum, pt, y, x, use_orientation, R)"]
    for i in range(32):
        terms = ["((SMOOTHED(%d, %d) < SMOOTHED(%d, %d)) << %d)" % (*table[8 * i + k], 7 - k) for k in range(8)]
        out.append("    desc[%d] = (uchar)(%s);" % (i, " + ".join(terms)))
    return "\n".join(out) + "\n#undef SMOOTHED\n"


if __name__ == "__main__":
    t = parse(open(sys.argv[1]).read())
    t.tofile(sys.argv[2])
    print("wrote %s: 256 tests, offsets in [%d, %d]" % (sys.argv[2], t.min(), t.max()))
