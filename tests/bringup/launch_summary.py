"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (+grid for GEMMs).  usage: launch_summary.py file.csv [topN]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in r:
    name, grid, t = row[4], row[8], float(row[-1])
    name = re.sub(r'\(.*', '', name).replace('void ', '').replace('uvc::', '')
    if 'gemm' in name: name += ' ' + grid
    agg[name][0] += 1; agg[name][1] += t; tot += t
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{v[1]/1e3:10.1f} us {v[0]:5d} {v[1]/v[0]/1e3:8.1f} us/launch {v[1]/tot*100:5.1f}%  {k[:100]}")
print(f"total {tot/1e3:.1f} us over {sum(v[0] for v in agg.values())} launches")
