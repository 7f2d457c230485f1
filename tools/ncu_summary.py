"""Summarise .ncu-rep files (read here, no GPU): one line per profiled launch with the counters the
roofline discussion needs.  Usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] > profiles/x.md"""
import csv
import io
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "dur_us",
    "dram__bytes_read.sum": "dram_rd_MB",
    "dram__bytes_write.sum": "dram_wr_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__cycles_active.avg": "cyc",
    "sm__cycles_elapsed.max": "cyc_max",
    "smsp__inst_executed_pipe_xu.sum": "xu_inst",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_conf",
    "launch__shared_mem_per_block_dynamic": "dsmem",
}


def rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(io.StringIO(out))
    hdr = next(rd)
    units = next(rd)
    for r in rd:
        yield dict(zip(hdr, r)), dict(zip(hdr, units))


def num(s):
    try:
        return float(s.replace(",", ""))
    except (ValueError, AttributeError):
        return None


for path in sys.argv[1:]:
    print(f"## {path}\n")
    first = True
    for r, u in rows(path):
        if first:
            avail = [k for k in r if any(k == w or k.startswith(w) for w in WANT)]
            first = False
        d = {}
        for k, name in WANT.items():
            v = num(r.get(k))
            if v is None:
                continue
            unit = u.get(k, "")
            if name == "dur_us":
                v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
            if name.endswith("_MB"):
                v = {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(unit, v)
            d[name] = v
        kn = r.get("Kernel Name", "?")[:40]
        print(f"- `{kn}` id={r.get('ID')} grid={r.get('launch__grid_size')} " +
              " ".join(f"{k}={v:.4g}" for k, v in d.items() if k not in ("grid",)))
    print()
