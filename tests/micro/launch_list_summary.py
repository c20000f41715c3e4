"""Groups an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid, block) and prints family shares.
Usage: python tests/micro/launch_list_summary.py launches.csv > by_kernel.txt"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ik, ig, ib, iv, iu = H.index("Kernel Name"), H.index("Grid Size"), H.index("Block Size"), H.index("Metric Value"), H.index("Metric Unit")
agg, total = {}, 0.0
for r in rows[hdr + 1:]:
    if len(r) <= iv or not r[iv]:
        continue
    us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    key = (re.sub(r"\(.*", "", r[ik])[:72], r[ig], r[ib])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
fam = {"gemm": 0.0, "bn": 0.0, "depthwise": 0.0, "other": 0.0}
for (k, _, _), (n, t) in agg.items():
    f = "gemm" if "gemm_tc" in k else "depthwise" if "dw3x3" in k else "bn" if re.search(r"bn_|bn3_|colstats", k) else "other"
    fam[f] += t
print("total us %.1f   (%d launches; cold-cache, serialised: shares, not absolutes)" % (total, sum(a[0] for a in agg.values())))
print("family shares: " + ", ".join("%s %.3f" % (k, v / total) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))
for (k, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %-14s %-13s n=%4d tot=%8.1f avg=%7.1f" % (k, g, b, n, t, t / n))
