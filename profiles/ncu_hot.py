"""Per-instruction stall breakdown of the first kernel in an `ncu --set full --import-source on` report.
    ncu -i X.ncu-rep --page source --csv > src.csv ; python profiles/ncu_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data, seen = [], set()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    a = r[idx['Address']]
    if a in seen:
        break
    seen.add(a); data.append(r)
I = lambda r, h: int(r[idx[h]] or 0)
tot = sum(I(r, '# Samples') for r in data)
inst = sum(I(r, 'Instructions Executed') for r in data)
print(len(data), "SASS instructions,", tot, "samples,", inst, "warp instructions executed")
stall_cols = [h for h in hdr if h.startswith('stall_') and '(Not Issued)' not in h]
agg = {h: sum(I(r, h) for r in data) for h in stall_cols}
print("stalls:", "  ".join("%s %.1f%%" % (h[6:], 100 * v / tot) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
for r in sorted(data, key=lambda r: -I(r, '# Samples'))[:top_n]:
    st = sorted(((h[6:], I(r, h)) for h in stall_cols), key=lambda kv: -kv[1])[:2]
    print("%6d %5.1f%%  %-72s %s" % (I(r, '# Samples'), 100 * I(r, '# Samples') / tot, r[idx['Source']].strip()[:72], st))
