#!/usr/bin/env python
"""Per-kernel DRAM traffic and headline counters out of `ncu --set full` reports.

    python tools/ncu_traffic.py gpurun_out/prof_a.ncu-rep [more.ncu-rep ...] > profiles/r01_ncu_traffic.json

Reads each report with `ncu -i <rep> --page raw --csv` (no GPU needed) and writes, per kernel name, the mean over
the captured launches of: duration, dram__bytes_read.sum + dram__bytes_write.sum (= roofline.traffic in bench.py),
DRAM / L2 / tensor-pipe / issue utilisation, registers, grid.  bench.py looks kernels up by substring."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_insts",
    "sm__cycles_elapsed.max": "cycles",
    # tensor-pipe utilisation (the north star asks for it beside the GEMMs) and the L2 / shared-memory pressure
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_of_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_of_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_lsu_wavefront_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_ghz",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    out = {}
    for rep in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        kn = hdr.index("Kernel Name")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").strip()
            d = out.setdefault(name, {"launches": 0, "report": rep.split("/")[-1]})
            d["launches"] += 1
            for i, h in enumerate(hdr):
                if h in KEEP and r[i] not in ("", "n/a"):
                    v = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                    d[KEEP[h]] = d.get(KEEP[h], 0.0) + v
    for d in out.values():
        n = d["launches"]
        for k in list(d):
            if k not in ("launches", "report"):
                d[k] = d[k] / n
        if "dram_read_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
    json.dump(out, sys.stdout, indent=1, sort_keys=True)
    print()


if __name__ == "__main__":
    main()
