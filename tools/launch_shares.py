"""Summarise an ncu launch list (csv from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file f`)
into per-kernel time shares.  Usage: python tools/launch_shares.py gpurun_out/launches.csv [skip_first_n_launches]"""
import csv, sys, collections, re

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = [r for r in rows if r is not hdr and len(r) > vi and r[hdr.index("Metric Name")] == "gpu__time_duration.sum"]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
data = data[skip:]
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in data:
    name = re.sub(r"\(.*", "", r[ki]).strip()
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
print(f"# launches {len(data)}, sum {s:.1f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {v:8.2f} ms {100 * v / s:5.1f}%  n={cnt[k]:4d}  {k[:90]}")
