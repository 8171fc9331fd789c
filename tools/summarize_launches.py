"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if r and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Metric Unit, Metric Value
agg = {}
for r in rows:
    if "gpu__time_duration.sum" not in r:
        continue
    name, unit, val = r[4], r[-2], float(r[-1].replace(",", ""))
    us = val / 1e3 if unit.startswith("ns") else (val * 1e3 if unit.startswith("ms") else val)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("tvt::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
print(f"# {sum(v[0] for v in agg.values())} launches captured, total {tot / 1e3:.2f} ms; cold-cache serialised times: compare SHARES")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k[:100]:100s} launches={v[0]:4d} us={v[1]:10.1f} share={v[1] / tot:6.3f}")
