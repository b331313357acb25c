"""One line per profiled launch out of an `ncu --page raw --csv` dump: duration, DRAM bytes and GB/s, tensor-pipe %, L2 / L1 throughput %, issue
utilisation, registers, achieved occupancy.  usage: python tools/ncu_table.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try: return float(r[idx[k]].replace(",", ""))
    except Exception: return d
print(f"{'kernel':58s} {'grid':>10s} {'us':>8s} {'rdMB':>7s} {'wrMB':>7s} {'GB/s':>6s} {'dram%':>6s} {'L2%':>5s} {'L1%':>5s} {'tens%':>6s} {'issue%':>6s} {'regs':>4s} {'occ%':>5s}")
for r in rows[2:]:
    name = r[idx['Kernel Name']].replace('void ', '').replace('uvc::', '').replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
    name = name.split('(')[0][:58]
    us = g(r, 'gpu__time_duration.sum')
    unit = rows[1][idx['gpu__time_duration.sum']]
    if unit == 'ns': us /= 1e3
    elif unit == 'ms': us *= 1e3
    rd, wr = g(r, 'dram__bytes_read.sum'), g(r, 'dram__bytes_write.sum')
    for k, v in (('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr')):
        u = rows[1][idx[k]]
        f = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
        if v == 'rd': rd *= f
        else: wr *= f
    gbs = (rd + wr) / us * 1e3 if us else 0     # MB / us = TB/s -> GB/s
    print(f"{name:58s} {r[idx['Grid Size']].replace(' ', ''):>10s} {us:8.1f} {rd:7.1f} {wr:7.1f} {gbs:6.0f} {g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} {g(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"{g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):6.1f} {g(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active', g(r, 'smsp__issue_active.avg.pct')):6.1f} "
          f"{int(g(r, 'launch__registers_per_thread')):4d} {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}")
