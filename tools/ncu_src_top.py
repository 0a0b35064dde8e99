"""Summarise the source page of an ncu report: top SASS instructions by stall samples.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv ; python tools/ncu_src_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_")]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    data.append((s, r))
tot = sum(s for s, _ in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for s, r in data:
    for h in stalls:
        try:
            agg[h] = agg.get(h, 0) + int(r[ix[h]] or 0)
        except ValueError:
            pass
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
exe = sum(int(r[ix["Instructions Executed"]] or 0) for _, r in data)
print("warp instructions executed:", exe)
for s, r in sorted(data, key=lambda x: -x[0])[:n]:
    top = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print("%6d %5.1f%%  %-70s %s exec=%s" % (s, 100.0 * s / max(tot, 1), r[ix["Source"]][:70],
                                            ",".join("%s=%d" % (h[6:], v) for v, h in top if v), r[ix["Instructions Executed"]]))
