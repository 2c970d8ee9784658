"""ncu launch list (gpu__time_duration.sum csv) -> per-step kernel shares.  python scratch/launch_summary.py csv first_timed_step n_steps"""
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
ui = h.index("Metric Unit")
launches = []
for r in rows[hdr + 1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    unit = r[ui]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    launches.append((r[ki], ms))
# a step starts at the gather kernel
starts = [i for i, (n, _) in enumerate(launches) if ("gather_rows_split" in n or "cov_fused_kernel" in n)]
print("%d launches, %d steps" % (len(launches), len(starts)))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for s in range(first, len(starts)):
    a = starts[s]
    b = starts[s + 1] if s + 1 < len(starts) else min(len(launches), a + (starts[s] - starts[s - 1]))
    agg = collections.OrderedDict()
    for n, ms in launches[a:b]:
        n = n.replace("void ", "").replace("onmf::", "")
        n = n[:n.index("(")] if "(" in n else n
        agg[n] = agg.get(n, 0.0) + ms
    tot = sum(agg.values())
    lars = sum(v for k, v in agg.items() if "lars_" in k)
    print("\n## step %d (timed step %d): %.2f ms of kernel time, LARS coder share %.1f%%\n" % (s + 1, s - first + 1, tot, 100 * lars / tot))
    print("| kernel | ms | share |\n|---|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1]):
        print("| `%s` | %.3f | %.1f%% |" % (k[:90], v, 100 * v / tot))
