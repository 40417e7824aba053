"""Per-kernel durations of a tracked frame in an ncu launch list of tools/frame_step_profile.py: mean over the tracked
frames (the first two frames are skipped) and the last frame.
   python tools/launch_frame.py launches.csv"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = [(r[ki].split("(")[0].split("::")[-1], float(r[vi].replace(",", "")) / (1000 if r[ui].startswith("n") else 1))
       for r in rows[start + 2:] if len(r) > vi]
first = [i for i, (n, _) in enumerate(seq) if "repitch" in n]
frames = [seq[a:b] for a, b in zip(first[2:], first[3:] + [len(seq)])]
frames = [f for f in frames if len(f) == len(frames[0])]
mean = OrderedDict()
for f in frames:
    for n, v in f:
        mean[n] = mean.get(n, 0.0) + v / len(frames)
last = dict(frames[-1])
print(f"{'kernel':45s} {'mean us':>8s} {'last':>8s}   ({len(frames)} frames)")
for n, v in mean.items():
    print(f"{n:45s} {v:8.2f} {last[n]:8.2f}")
print(f"{'total':45s} {sum(mean.values()):8.2f} {sum(last.values()):8.2f}")
