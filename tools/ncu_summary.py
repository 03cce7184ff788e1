#!/usr/bin/env python
"""Extracts the numbers DESIGN.md / bench.py quote from an `ncu --set full` report (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep <tag> [--update-latest]
writes profiles/r2/ncu_<tag>_summary.json (+ the raw page restricted to the quoted metrics as csv) and, with
--update-latest, merges the kernels into profiles/ncu_full_latest_summary.json under "<tag> <kernel name>"."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["Kernel Name", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    out = {}
    os.makedirs(os.path.join(ROOT, "profiles", "r2"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r2", f"ncu_{tag}_raw.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in cols]); w.writerow([units[c] for c in cols])
        for r in rows[2:]:
            w.writerow([r[c] for c in cols])
            name = re.sub(r"^void |kgeb::tcb::|kgeb::tc::|\(.*$", "", r[hdr.index("Kernel Name")]).strip()
            d = {}
            for c in cols[1:]:
                try:
                    d[hdr[c]] = float(r[c].replace(",", "")) * (UNIT.get(units[c], 1.0) if hdr[c].split(".")[0] in
                                                                   ("gpu__time_duration", "dram__bytes_read", "dram__bytes_write", "lts__t_bytes") else 1.0)
                except ValueError:
                    pass
            d["dram_bytes_read"], d["dram_bytes_write"] = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
            d["duration_s"] = d.get("gpu__time_duration.sum")
            if "sm__cycles_active.avg" in d and d.get("sm__cycles_elapsed.avg"):
                d["sm_active_over_elapsed"] = d["sm__cycles_active.avg"] / d["sm__cycles_elapsed.avg"]
            out[f"{tag} {name}"] = d
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2", f"ncu_{tag}_summary.json"), "w"), indent=1)
    if "--update-latest" in sys.argv:
        path = os.path.join(ROOT, "profiles", "ncu_full_latest_summary.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur.update(out)
        json.dump(cur, open(path, "w"), indent=1)
    for k, d in out.items():
        print(k, {x: (round(v, 4) if isinstance(v, float) else v) for x, v in d.items() if x in
                  ("duration_s", "dram_bytes_read", "dram_bytes_write", "sm_active_over_elapsed",
                   "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                   "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")})


if __name__ == "__main__":
    main()
