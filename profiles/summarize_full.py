"""Summarise an `ncu --set full` raw-page CSV of one pipeline step (tools/profile_round.sh): per kernel the launch count,
device time, DRAM bytes read/written, achieved HBM GB/s (dram bytes / duration), and the utilisation percentages
(DRAM, L2, SM, tensor pipe).  Usage: python profiles/summarize_full.py <raw.csv> <out.csv> "<title>" [traffic.json]"""
import collections
import csv
import json
import re
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[hi], rows[hi + 1]
col = {n: i for i, n in enumerate(hdr)}
M = {"t": "gpu__time_duration.sum", "r": "dram__bytes_read.sum", "w": "dram__bytes_write.sum",
     "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
     "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "regs": "launch__registers_per_thread", "xbar": "l1tex__m_xbar2l1tex_read_bytes.sum"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def val(r, key):
    c = col.get(M[key])
    if c is None:
        return 0.0
    try:
        v = float(r[c].replace(",", ""))
    except ValueError:
        return 0.0
    return v * SCALE.get(units[c], 1.0)


agg = collections.OrderedDict()
for r in rows[hi + 2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", re.sub(r"^void ", "", r[col["Kernel Name"]])).replace("<unnamed>::", "")
    if name.startswith("at::"):
        continue
    d = agg.setdefault(name, collections.defaultdict(float))
    d["n"] += 1
    for k in M:
        d[k] += val(r, k)
out = ["# " + title,
       "# ncu --set full --clock-control none, one pipeline step (cold cache, ~40 replays per launch: compare shares, not absolutes)",
       "# hbm_gbs = (dram read + write bytes) / device time; percentages are averages over the kernel's launches",
       "kernel,launches,total_us,dram_read_MB,dram_write_MB,hbm_gbs,dram_pct,l2_pct,sm_pct,tensor_pct,l2_to_sm_read_MB,regs"]
tot = sum(d["t"] for d in agg.values())
for k, d in agg.items():
    n = d["n"]
    gbs = (d["r"] + d["w"]) / (d["t"] * 1e-6) / 1e9 if d["t"] > 0 else 0.0
    out.append("%s,%d,%.1f,%.1f,%.1f,%.0f,%.1f,%.1f,%.1f,%.1f,%.1f,%d" % (
        k, n, d["t"], d["r"] / 1e6, d["w"] / 1e6, gbs, d["dram"] / n, d["l2"] / n, d["sm"] / n, d["tensor"] / n,
        d["xbar"] / 1e6, d["regs"] / n))
out.append("# total device time of the step under ncu: %.1f us" % tot)
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
if len(sys.argv) > 4:
    frames = int(sys.argv[5]) if len(sys.argv) > 5 else 32
    tj = {"frames_in_capture": frames, "source": src}
    for k, d in agg.items():
        tj[k.split("<")[0] if k.startswith("k_conv3x3") else k] = {
            "dram_bytes_per_launch": (d["r"] + d["w"]) / d["n"], "launches": int(d["n"]), "us_per_launch": d["t"] / d["n"]}
    json.dump(tj, open(sys.argv[4], "w"), indent=1)
