"""Compare two timelines written by tools/step_timeline.py (e.g. 1 GPU vs N GPUs on the same box): per kernel name, the
time per step in each and the difference; per stream busy time; compute-stream idle time."""
import json, sys
a, b = (json.load(open(p)) for p in sys.argv[1:3])
def summarise(t):
    # drop the first profiled step (the ranks' profilers start at different times: one long wait at the first barrier)
    rows = t["rows"]
    ends = [i for i, r in enumerate(rows) if "adamw_ema" in r["name"]]
    per = len(ends) // t["nstep"]
    if t["nstep"] > 1 and per:
        rows = rows[ends[per - 1] + 1:]
        t = dict(t, rows=rows, nstep=t["nstep"] - 1)
    n = t["nstep"]; by = {}; streams = {}
    for r in t["rows"]:
        k = r["name"].split("(")[0][:60]
        d = by.setdefault(k, [0, 0.0]); d[0] += 1 / n; d[1] += r["dur"] / 1e3 / n
        streams.setdefault(r["stream"], []).append(r)
    main = max(streams, key=lambda s: sum(r["dur"] for r in streams[s]))
    rs = streams[main]
    idle = sum(max(0, y["ts"] - (x["ts"] + x["dur"])) for x, y in zip(rs, rs[1:])) / 1e3 / n
    span = (max(r["ts"] + r["dur"] for r in t["rows"]) - t["rows"][0]["ts"]) / 1e3 / n
    return by, idle, span, sum(r["dur"] for r in rs) / 1e3 / n
A, ia, sa, ba = summarise(a); Bb, ib, sb, bb = summarise(b)
print(f"world {a['world']} -> {b['world']}: step {sa:.2f} -> {sb:.2f} ms (under CUPTI); compute stream busy {ba:.2f} -> {bb:.2f}, idle {ia:.2f} -> {ib:.2f}")
keys = sorted(set(A) | set(Bb), key=lambda k: -abs(Bb.get(k, [0, 0])[1] - A.get(k, [0, 0])[1]))
print(f"{'kernel':60s} {'n':>6s} {'ms/step':>9s} -> {'n':>6s} {'ms/step':>9s}  {'diff':>7s}")
for k in keys[:30]:
    x, y = A.get(k, [0, 0.0]), Bb.get(k, [0, 0.0])
    print(f"{k:60s} {x[0]:6.0f} {x[1]:9.2f} -> {y[0]:6.0f} {y[1]:9.2f}  {y[1] - x[1]:+7.2f}")
