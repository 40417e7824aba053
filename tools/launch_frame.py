"""Per-kernel durations of the last complete frame in an ncu launch list of tools/frame_step_profile.py.
   python tools/launch_frame.py launches.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = [(r[ki].split("(")[0].split("::")[-1], float(r[vi].replace(",", "")) / (1000 if r[ui].startswith("n") else 1))
       for r in rows[start + 2:] if len(r) > vi]
first = [i for i, (n, _) in enumerate(seq) if "repitch" in n]
frame = seq[first[-2]:first[-1]]
for n, v in frame:
    print(f"{n:45s} {v:8.2f} us")
print(f"{'total':45s} {sum(v for _, v in frame):8.2f} us")
