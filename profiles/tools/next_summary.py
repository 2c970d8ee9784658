"""Markdown table from `bench.py --workload next` logs: python profiles/tools/next_summary.py BEFORE.log AFTER.log > profiles/r2_next_rows.md"""
import json
import sys


def rows(path):
    out = {}
    for line in open(path):
        if not line.startswith("{"):
            continue
        r = json.loads(line)
        key = (r["kernel"], r["shape"])
        if key in out and "ms_median" not in r:
            out[key].update(r)
        else:
            out.setdefault(key, {}).update(r)
    return out


before, after = rows(sys.argv[1]), rows(sys.argv[2])
print("| kernel | shape | ms before | ms now | algorithmic MB | GB/s | of HBM copy peak | vs write-only fill | unit rate | CPU form (1 core) | GPU / CPU |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
fills = {k[1]: v for k, v in after.items() if "write_only_fill_ms" in v and "ms_median" not in v}
for key, r in after.items():
    if "ms_median" not in r and "wall_ms_host_to_host" not in r:
        continue
    b = before.get(key, {})
    ms, msb = r.get("ms_median", r.get("wall_ms_host_to_host")), b.get("ms_median", b.get("wall_ms_host_to_host"))
    unit = [(k, v) for k, v in r.items() if k.startswith("M_")]
    cpu = [(k, v) for k, v in r.items() if k.startswith("cpu_") and k.endswith("/s")]
    fill = r.get("frac_of_fill_rate")
    if fill is None:
        for tag, f in fills.items():
            if key[1].startswith(tag):
                fill = f["frac_of_fill_rate"]
    print("| `%s` | %s | %s | %.3f%s | %s | %s | %s | %s | %s | %s | %s |" % (
        key[0], key[1], ("%.3f" % msb) if msb else "—", ms, " (wall, host to host)" if "wall_ms_host_to_host" in r else "",
        ("%.0f" % r["algorithmic_MB"]) if "algorithmic_MB" in r else "—",
        ("%.0f" % r["GB/s"]) if "GB/s" in r else "—",
        ("%.2f" % r["frac_hbm_peak"]) if "frac_hbm_peak" in r else "—",
        ("%.2f" % fill) if fill else "—",
        ("%.1f %s" % (unit[0][1], unit[0][0])) if unit else (("%.2f us/atom" % r["us_per_atom"]) if "us_per_atom" in r else "—"),
        ("%.0f /s" % cpu[0][1]) if cpu else (("%.2f ms" % r["cpu_ms"]) if "cpu_ms" in r else "—"),
        ("%.0f x" % r["gpu_over_cpu"]) if "gpu_over_cpu" in r else "—"))
