#!/usr/bin/env python
"""Condenses an Nsight Compute report (ncu -i X.ncu-rep --page raw --csv) into the per-kernel table kept
under profiles/.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/NAME.md"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_inst"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "math_pipe_throttle", "barrier", "wait", "not_selected", "selected",
          "lg_throttle", "mio_throttle", "dispatch_stall", "branch_resolving", "no_instructions", "membar", "drain", "sleeping"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{path}` (`--set full --clock-control none`; per-launch, cold-cache, serialised)\n")
    print("| kernel | " + " | ".join(n for _, n in KEYS) + " | top stalls (pc samples) |")
    print("|---|" + "---|" * (len(KEYS) + 1))
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("fpc::", "")
        vals = []
        for k, n in KEYS:
            v = r[col[k]] if k in col else ""
            try:
                f = float(v.replace(",", ""))
                u = units[col[k]] if k in col else ""
                if u == "ns":
                    f /= 1e3
                if u == "byte":
                    f /= 1e6
                if u == "Kbyte":
                    f /= 1e3
                if u == "Gbyte":
                    f *= 1e3
                v = f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}"
            except ValueError:
                pass
            vals.append(v)
        st = []
        for s in STALLS:
            k = f"smsp__pcsamp_warps_issue_stalled_{s}"
            if k in col:
                try:
                    st.append((float(r[col[k]].replace(",", "")), s))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        top = ", ".join(f"{s} {v / tot * 100:.0f}%" for v, s in sorted(st, reverse=True)[:4])
        print(f"| {name} | " + " | ".join(vals) + f" | {top} |")


if __name__ == "__main__":
    main(sys.argv[1])
