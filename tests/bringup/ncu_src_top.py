"""Summarise an `ncu --page source --csv` dump: top stalled SASS instructions with their dominant stall reasons.
usage: ncu -i X.ncu-rep --page source --csv --kernel-id ::regex:NAME:N > src.csv ; python ncu_src_top.py src.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; idx = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
def f(x):
    try: return float(x)
    except Exception: return 0.0
tot = sum(f(r[idx['# Samples']]) for r in data)
print("kernel:", rows[0][1] if rows[0] else "?", "| total samples", tot, "| instructions", len(data))
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
agg = {k: sum(f(r[idx[k]]) for r in data) for k in stalls}
print("stall totals:", ", ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, r in sorted(enumerate(data), key=lambda t: -f(t[1][idx['# Samples']]))[:topn]:
    s = f(r[idx['# Samples']])
    st = sorted(((f(r[idx[k]]), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{n:5d} {s:7.0f} {s/tot*100:5.1f}% {r[idx['Source']][:100]:100s} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}")
