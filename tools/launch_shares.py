"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name.
usage: python tools/launch_shares.py launches.csv [n_steps_in_capture]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= iv: continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    name = r[ik].split("(")[0][:60]
    agg[name] += v; cnt[name] += 1
tot = sum(agg.values())
print(f"# launches {sum(cnt.values())}, sum {tot:.1f} ms ({tot / steps:.1f} ms per step over {steps:g} steps)")
for k, v in agg.most_common(40):
    print(f"{v / steps:9.2f} ms {100 * v / tot:5.1f}%  n={cnt[k] / steps:7.1f}  {k}")
