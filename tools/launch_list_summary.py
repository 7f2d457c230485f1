"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv`) and the `profiles/traffic.json` that bench.py reads for `roofline.traffic`.
Usage: python tools/launch_list_summary.py gpurun_out/launches_TAG.csv TAG   (writes profiles/TAG_*.{csv,md}, traffic.json)"""
import csv
import json
import os
import re
import shutil
import sys
from collections import OrderedDict

src, tag = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = [l for l in open(src) if l.startswith('"')]
rd = csv.DictReader(lines)
launches = OrderedDict()
for r in rd:
    d = launches.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)   # -> ms
    elif unit in ("Kbyte", "Mbyte", "Gbyte"):
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    d[r["Metric Name"]] = v


def short(name):
    m = re.search(r"(\w*kernel\w*)\s*[<(]", name)
    base = m.group(1) if m else name[:32]
    t = re.search(base + r"<\(int\)(\d+)", name)
    if t:
        return f"{base}<{t.group(1)}>"
    t = re.search(base + r"<\(bool\)(\d)>", name)
    if t and base == "conv_gemm_kernel":
        return f"{base}<{'pair' if t.group(1) == '1' else 'single'}>"
    return base


agg = OrderedDict()
for d in launches.values():
    a = agg.setdefault(short(d["name"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list of one warm UNet step (UNet batch 32), build {tag}: {len(launches)} launches, {tot:.2f} ms summed "
       "(cold-cache, serialised)", "| kernel | launches | ms | share | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e6:.0f} | {a[3] / 1e6:.0f} |")
os.makedirs(os.path.join(root, "profiles"), exist_ok=True)
open(os.path.join(root, "profiles", f"{tag}_launch_list_summary.md"), "w").write("\n".join(out) + "\n")
shutil.copy(src, os.path.join(root, "profiles", f"{tag}_launches.csv"))
gs = [a for k, a in agg.items() if k.startswith("conv_gemm_kernel")]
g = [sum(a[i] for a in gs) for i in range(4)] if gs else None
if g:
    json.dump({"conv_gemm_kernel": {"launches": g[0], "dram_read_bytes": g[2], "dram_write_bytes": g[3],
                                    "source": f"profiles/{tag}_launches.csv (ncu, one UNet step at UNet batch 32, cold cache per launch)"}},
              open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(out))
