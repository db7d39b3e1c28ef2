"""Turns an ncu --csv launch list (gpu__time_duration.sum [+ dram bytes]) into a per-kernel markdown table of shares."""
import collections
import csv
import re
import sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = [r for r in csv.reader(open(path, errors="replace")) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
hdr = rows[hdr_i]
ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
per = collections.defaultdict(lambda: {"n": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}
for r in rows[hdr_i + 1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void ", "", name)
    try:
        v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    except ValueError:
        continue
    d = per[name]
    d["n"].add(r[iid])
    if r[im].startswith("gpu__time_duration"):
        d["ns"] += v
    elif r[im].startswith("dram__bytes_read"):
        d["rd"] += v
    elif r[im].startswith("dram__bytes_write"):
        d["wr"] += v
tot = sum(d["ns"] for d in per.values())
print(f"# {title}")
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
print("| kernel | launches | total ms | share | DRAM read GB | DRAM write GB |")
print("|---|---:|---:|---:|---:|---:|")
for name, d in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
    print(f"| `{name}` | {len(d['n'])} | {d['ns'] / 1e6:.3f} | {100 * d['ns'] / tot:.1f}% | {d['rd'] / 1e9:.3f} | {d['wr'] / 1e9:.3f} |")
print(f"| **total** | {sum(len(d['n']) for d in per.values())} | {tot / 1e6:.3f} | 100% | {sum(d['rd'] for d in per.values()) / 1e9:.2f} | {sum(d['wr'] for d in per.values()) / 1e9:.2f} |")

if len(sys.argv) > 3:      # third argument: write the mean DRAM bytes per launch of the kernels matching argv[4] (default umma)
    import json
    pat = sys.argv[4] if len(sys.argv) > 4 else "umma"
    sel = [d for n, d in per.items() if pat in n]
    launches = sum(len(d["n"]) for d in sel)
    total = sum(d["rd"] + d["wr"] for d in sel)
    json.dump({"kernel_pattern": pat, "launches": launches, "dram_bytes_total": total, "dram_bytes_per_launch": total / max(launches, 1),
               "source": path}, open(sys.argv[3], "w"), indent=1)
