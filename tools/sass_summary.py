#!/usr/bin/env python3
"""Per-kernel SASS mnemonic counts of libvslam_b200.so (cuobjdump -sass): the evidence that the kernels are compiled for
sm_100a and use what DESIGN.md says they use -- TMA box loads (UTMALDG) with mbarrier transaction counts (SYNCS),
packed fp32 (FFMA2 / FADD2), packed 16-bit integer min/max (VIMNMX3), byte SIMD (VABSDIFF4), POPC, warp reductions
(REDUX), FP64 (DFMA / DADD / DMUL / MUFU.RCP64H for the IEEE divisions).  No UTC*MMA / TMEM: nothing is a contraction.
Usage: tools/sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vslam-pose-estimation-framework_b200", "libvslam_b200.so")
WATCH = ["UTMALDG", "SYNCS", "UTCMMA", "UTCHMMA", "HMMA", "FFMA2", "FADD2", "FMUL2", "FFMA", "VIMNMX3", "VIMNMX", "VABSDIFF4", "POPC",
         "REDUX", "SHFL", "VOTE", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "BAR",
         "PRMT", "SHF", "LOP3", "IMAD", "BREV", "FLO"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            short = re.sub(r"\(.*", "", name).replace("vslam::", "").replace("(anonymous namespace)::", "")
            short = re.sub(r"^void\s+", "", short).strip() or m.group(1)
            g = re.search(r"_cu_[0-9a-f]{8}(\d+)", short)     # anonymous-namespace kernels: c++filt does not know nvcc's TU hash
            if g:
                start = g.end()
                n = int(g.group(1))
                t = re.match(r"ILi(\d+)EE", short[start + n:])
                short = short[start:start + n] + ("<%s>" % t.group(1) if t else "")
            cur = kernels.setdefault(short, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            cur[m.group(1)] += 1
            if m.group(1) in ("MUFU", "VIMNMX3", "VIMNMX", "LDS", "LDG", "STG", "SYNCS"):
                cur[m.group(1) + m.group(2)] += 1
    print("# %s: arch %s, %d kernels (static SASS instruction counts, not executed counts)" % (os.path.basename(LIB), ", ".join(arch), len(kernels)))
    for name, c in kernels.items():
        hits = ["%s %d" % (w, c[w]) for w in WATCH if c[w]]
        detail = ["%s %d" % (k, v) for k, v in sorted(c.items()) if "." in k and any(k.startswith(p) for p in ("MUFU", "VIMNMX3", "SYNCS", "LDS.", "LDG."))]
        print("\n%s  [%d instructions]\n  %s" % (name, c["_total"], "  ".join(hits)))
        if detail:
            print("  " + "  ".join(detail))


if __name__ == "__main__":
    main()
