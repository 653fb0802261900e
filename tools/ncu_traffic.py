#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` capture (run HERE, no GPU needed):

    python tools/ncu_traffic.py gpurun_out/r02_path_full.ncu-rep cfg2_b32 [--md profiles/r02_ncu_full_path.md]

Writes / updates profiles/r02_ncu_traffic.json ({"<workload>_b<frames>": {"<kernel>": dram bytes per launch}}, read by
bench.py for `roofline.traffic`) and optionally a markdown table with duration, DRAM bytes and throughput, pipe
utilisation, issue activity, registers and occupancy of every kernel in the capture (averaged over its launches)."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
        "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
COLS = {
    "dur_us": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "fma_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "alu_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "occ_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "inst": "smsp__inst_executed.sum",
}


def main():
    rep, key = sys.argv[1], sys.argv[2]
    md = sys.argv[sys.argv.index("--md") + 1] if "--md" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    idx = {k: hdr.index(v) for k, v in COLS.items() if v in hdr}
    agg = {}
    for r in rows[2:]:
        name = re.sub(r"^(void )?(fpc::)?", "", r[ik]).split("<")[0].split("(")[0]
        name = re.sub(r"^(k_argmax_runs|k_emit_runs)(_[a-z]+\d+|\d+)$", r"\1", name)     # variants report under bench.py's kernel name
        a = agg.setdefault(name, {"n": 0, **{k: 0.0 for k in idx}})
        a["n"] += 1
        for k, i in idx.items():
            v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0.0
            a[k] += v * UNIT.get(units[i], 1.0)
    for a in agg.values():
        for k in idx:
            a[k] /= a["n"]
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[key] = {k: round(a["rd"] + a["wr"]) for k, a in agg.items()}
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    if md:
        with open(md, "w") as f:
            f.write(f"# ncu --set full, {key} ({os.path.basename(rep)}); per launch, averaged over the captured launches\n\n")
            f.write("ncu serialises kernels and replays each ~40 times with cold caches: compare SHARES, not absolute times.\n\n")
            f.write("| kernel | launches | time us | dram read MB | dram write MB | dram % | fma pipe % | alu pipe % | issue % | regs | warps active % | warp instr |\n")
            f.write("|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            tot = sum(a["dur_us"] for a in agg.values())
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["dur_us"]):
                f.write(f"| {k} | {a['n']} | {a['dur_us']:.1f} ({100 * a['dur_us'] / tot:.0f} %) | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | "
                        f"{a.get('dram_pct', 0):.0f} | {a.get('fma_pct', 0):.0f} | {a.get('alu_pct', 0):.0f} | {a.get('issue_pct', 0):.0f} | "
                        f"{a.get('regs', 0):.0f} | {a.get('occ_pct', 0):.0f} | {a.get('inst', 0):.3g} |\n")
    print(json.dumps(table[key], indent=1))


if __name__ == "__main__":
    main()
