"""Top stall sites of one profiled launch from `ncu --page source --csv` (SASS view).
Usage: python tools/ncu_source_top.py report.ncu-rep [launch_index] [top_n]"""
import csv
import subprocess
import sys
from collections import Counter

path = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout.splitlines()
# the csv holds one block per profiled launch: a "Kernel Name" line, a header, then SASS rows
blocks, cur = [], None
for line in out:
    if line.startswith('"Kernel Name"'):
        cur = [line]
        blocks.append(cur)
    elif cur is not None:
        cur.append(line)
blk = blocks[idx if len(blocks) > idx else 0]
print(blk[0][:200], f"(block {idx} of {len(blocks)})")
rd = csv.reader(blk[1:])
hdr = next(rd)
rows = [dict(zip(hdr, r)) for r in rd if len(r) == len(hdr) and r[0] != hdr[0]]
tot = sum(int(r["# Samples"] or 0) for r in rows)
texec = sum(int(r["Instructions Executed"] or 0) for r in rows)
print(f"instructions: {len(rows)} SASS lines, {texec} warp-instr executed, {tot} samples")
reasons = Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows:
    for c in stall_cols:
        reasons[c] += int(r[c] or 0)
print("stall totals:", ", ".join(f"{k[6:]}={v} ({100 * v / max(tot, 1):.1f}%)" for k, v in reasons.most_common(10)))
mix = Counter()
for r in rows:
    op = r["Source"].split()
    op = [o for o in op if not o.startswith("@")]
    mix[op[0].split(".")[0] if op else "?"] += int(r["Instructions Executed"] or 0)
print("instr mix:", ", ".join(f"{k}={100 * v / max(texec, 1):.1f}%" for k, v in mix.most_common(16)))
print(f"--- top {top} SASS lines by samples")
for i, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"] or 0))[:top]:
    s = int(r["# Samples"] or 0)
    why = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{100 * s / max(tot, 1):5.1f}%  #{i:<5d} exec={r['Instructions Executed']:>9s}  {r['Source'].strip()[:70]:70s} {why}")
