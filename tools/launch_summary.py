"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv [skip_first_n_launches]
Splits the list into steps at the given marker kernel substring (optional 3rd argument) and summarises the LAST step."""
import collections, csv, re, sys
path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[start + 1:] if len(r) > vi and r[mi] == "gpu__time_duration.sum"]
per_step = int(sys.argv[3]) if len(sys.argv) > 3 else 1       # marker launches per step
if marker:
    marks = [i for i, (k, _) in enumerate(data) if marker in k]
    # one full step: from the first marker of the last-but-one step to the first marker of the last step
    a, b = marks[-2 * per_step], marks[-per_step]
    data = data[a:b]
tot = sum(v for _, v in data)
agg = collections.OrderedDict()
for k, v in data:
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", k))[:100]
    e = agg.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += v
print(f"{len(data)} launches, {tot / 1e3:.1f} us of kernel time (cold-cache, serialised)")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e3:9.1f} us {100 * t / tot:5.1f} %  x{n:<4d} {name}")
