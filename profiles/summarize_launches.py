"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares for ONE
pipeline pass (the last `k_insert` .. end).  Usage: python profiles/summarize_launches.py <csv> <out.csv> "<title>" """
import collections
import csv
import re
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(src) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
rows = []
for row in r:
    if len(row) <= vi:
        continue
    try:
        v = float(row[vi].replace(",", ""))
    except ValueError:
        continue
    u = row[ui]
    ns = v * 1e3 if u in ("us", "usecond") else v * 1e6 if u in ("ms", "msecond") else v
    rows.append((row[ki], ns))
starts = [i for i, (n, _) in enumerate(rows) if "k_insert" in n]
passes = len(starts)
s = starts[-1]
step = rows[s:]
agg = collections.OrderedDict()
seq = []
for n, ns in step:
    short = re.sub(r"^void ", "", n)
    short = re.sub(r"\(.*", "", short)
    short = re.sub(r"<unnamed>::", "", short)[:70]
    agg.setdefault(short, [0.0, 0])
    agg[short][0] += ns
    agg[short][1] += 1
    seq.append((short, ns))
tot = sum(v[0] for v in agg.values())
out = ["# " + title, "# gpu__time_duration.sum per launch, --clock-control none (cold-cache, serialised: compare SHARES)",
       "# %d pipeline passes captured; this is the last one: %.3f ms over %d launches" % (passes, tot / 1e6, len(step)),
       "kernel,launches,total_us,share"]
for k, (ns, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.append("%s,%d,%.1f,%.4f" % (k, c, ns / 1e3, ns / tot))
out.append("# launch sequence (us): " + " | ".join("%s %.1f" % (k.split("<")[0][:18], ns / 1e3) for k, ns in seq))
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
