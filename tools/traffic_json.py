"""profiles/gemm2_traffic.json entry from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` capture of
the CTA-pair GEMM launches of whole bench steps.  usage: python tools/traffic_json.py capture.csv key "how it was captured" """
import csv, json, os, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    per.setdefault(r[idx["ID"]], {})[r[idx["Metric Name"]]] = (float(r[idx["Metric Value"]].replace(",", "")), r[idx["Metric Unit"]])
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
tot_b = tot_us = 0.0
for m in per.values():
    tot_b += sum(v * scale[u] for k, (v, u) in m.items() if k.startswith("dram__bytes"))
    v, u = m["gpu__time_duration.sum"]; tot_us += v * scale[u]
n = len(per)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "gemm2_traffic.json")
try:
    d = json.load(open(path))
    if "dram_bytes_per_launch" in d: d = {"small_s1:tf32": d}        # the round-1 record
except Exception:
    d = {}
d[sys.argv[2]] = {"dram_bytes_per_launch": round(tot_b / n), "launches": n, "avg_launch_us_under_ncu": round(tot_us / n, 2), "source": sys.argv[3]}
json.dump(d, open(path, "w"), indent=1)
print(sys.argv[2], d[sys.argv[2]])
