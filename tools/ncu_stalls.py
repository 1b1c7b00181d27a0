"""Summarise an `ncu --page source --csv --print-source sass` export: stall-reason totals and the hottest instructions.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv; python tools/ncu_stalls.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
# the export holds one block per kernel launch: "Kernel Name" line, header line, instruction lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
b = blocks[0]
idx = {h: i for i, h in enumerate(b["hdr"])}
stall_cols = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[idx["# Samples"]] or 0) for r in b["data"])
print(b["name"][:100], "| instructions", len(b["data"]), "| samples", tot)
agg = {h: sum(int(r[idx[h]] or 0) for r in b["data"]) for h in stall_cols}
print("stall totals:", ", ".join(f"{k[6:]} {v} ({100 * v / max(tot, 1):.0f}%)" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
for r in sorted(b["data"], key=lambda r: -int(r[idx["# Samples"]] or 0))[:topn]:
    st = sorted(((h[6:], int(r[idx[h]] or 0)) for h in stall_cols), key=lambda x: -x[1])[:2]
    print(r[idx["# Samples"]].rjust(7), r[idx["Source"]].strip()[:88].ljust(88), st)
