"""Summarise an .ncu-rep (one kernel) and an ncu launch-list csv into text for profiles/."""
import csv, subprocess, sys, collections, io

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__inst_executed_pipe_tensor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex.sum", "lts__t_bytes.sum", "sm__inst_executed_pipe_uniform", "smsp__warp_issue_stalled"]


def rep_summary(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"== kernel: {d.get('Kernel Name','?')}  grid {d.get('Grid Size','?')} block {d.get('Block Size','?')}")
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in KEYS) and "peak_sustained" not in h.replace("pct_of_peak_sustained", ""):
                lines.append(f"  {h} [{u}] = {v}")
    return "\n".join(lines)


def launches_summary(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            unit = d["Metric Unit"]
            v = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v * 1e3 if unit == "s" else v
            agg.setdefault(d["Kernel Name"][:90], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    return "\n".join(f"{k:92s} n={len(v):3d} total_ms={sum(v):10.3f} share={sum(v)/tot:.4f}" for k, v in agg.items())


if __name__ == "__main__":
    for a in sys.argv[1:]:
        print(f"#### {a}")
        print(rep_summary(a) if a.endswith(".ncu-rep") else launches_summary(a))
