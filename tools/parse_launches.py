"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-launch time of the last step."""
import csv
import sys

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[start + 1:] if r[mi] == "gpu__time_duration.sum"]
n = len(data) // steps
last = data[-n:]
tot = sum(v for _, v in last)
for k, v in last:
    print("%-100s %9.1f us %5.1f%%" % (k[:100], v / 1000, 100 * v / tot))
print("launches/step %d   total %.1f us" % (n, tot / 1000))
