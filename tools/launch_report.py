"""Developer tool: summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per frame."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches_f16.csv"
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))


def t_us(row):
    t = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return t / 1000 if u == "ns" else (t * 1000 if u == "ms" else t)


idx = [i for i, r in enumerate(rows) if "image_to_nhwc" in r["Kernel Name"]]
a, b = (idx[1], idx[2]) if len(idx) > 2 else (idx[0], idx[1])
tot = 0
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[a:b]:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("smot::", "")
    t = t_us(r)
    tot += t
    agg[n][0] += 1
    agg[n][1] += t
    if t > thr:
        print("%7.1f  grid=%-14s blk=%-12s %s" % (t, r["Grid Size"], r["Block Size"], n[:70]))
print("frame total %.1f us, %d launches" % (tot, b - a))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%8.1f us %4d  %s" % (v[1], v[0], k[:90]))
