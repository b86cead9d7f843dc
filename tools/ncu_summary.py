"""One table row per profiled launch from ncu --set full reports:  python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]
Columns: duration, DRAM bytes (read+write), L2 bytes, DRAM %, FP64-pipe %, DMMA instruction %, issue-slot %, registers,
top stall reasons (warp-cycles per issued instruction)."""
import csv, io, subprocess, sys

KEYS = {
    "dur_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "l2_bytes": "lts__t_bytes.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "fp64_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "dmma_pct": "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
}
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


print("| kernel | grid x block | regs | us | DRAM MB (rd+wr) | L2 MB | DRAM % | FP64 pipe % | DMMA % | issue % | top stalls (cycles/issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for rep in sys.argv[1:]:
    hdr, units, rows = rows_of(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        def val(key):
            i = idx.get(KEYS[key])
            if i is None or r[i] == "":
                return float("nan")
            return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        stalls = []
        for h, i in idx.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        name = r[idx["Kernel Name"]].split("(")[0]
        print(f"| {name} | {int(val('grid'))} x {int(val('block'))} | {int(val('regs'))} | {val('dur_us'):.1f} | "
              f"{(val('dram_rd') + val('dram_wr')) / 1e6:.2f} | {val('l2_bytes') / 1e6:.1f} | {val('dram_pct'):.1f} | {val('fp64_pct'):.1f} | "
              f"{val('dmma_pct'):.1f} | {val('issue_pct'):.1f} | " + ", ".join(f"{n} {v:.1f}" for v, n in stalls[:4]) + " |")
