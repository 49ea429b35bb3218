"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
    python profiles/summarize_launches.py profiles/r1_launches_ppo_fused.csv [top]
Per-launch times under ncu are cold-cache and serialised: read the SHARES, not the absolutes."""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 15
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ik, iv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 2:]:
    if len(r) <= iv:
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    k = r[ik][:70]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in agg.values())
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k:72s} n={n:5d} total={v / 1e3:10.1f} us  avg={v / n / 1e3:8.2f} us  {100 * v / tot:5.1f}%")
