#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of
numbers DESIGN.md / profiles/ quote.  Usage: tools/ncu_summary.py report.ncu-rep [frames]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_lookup_hit.sum", "l1tex__t_sectors_lookup_miss.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sass__inst_executed_global_loads", "sass__inst_executed_global_stores", "sass__inst_executed_shared_loads",
    "sass__inst_executed_shared_stores", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"), "| grid", d.get("launch__grid_size"))
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>16s} {u[k]}")
        st = {h[len(STALL):]: float(d[h]) for h in hdr if h.startswith(STALL) and not h.endswith("_not_issued")}
        tot = sum(st.values())
        print("  stall samples (pc sampling, all):")
        for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
            print(f"    {k:28s} {100 * v / tot:5.1f} %")
        if frames:
            sms = 148
            cyc = float(d["sm__cycles_elapsed.avg"])
            inst = float(d["smsp__inst_executed.sum"])
            wf = float(d["l1tex__data_pipe_lsu_wavefronts.avg"]) * sms if "l1tex__data_pipe_lsu_wavefronts.avg" in d else \
                float(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]) / 100 * cyc * sms
            wfs = float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"])
            print(f"  per frame per SM: {cyc * sms / frames:7.1f} cycles, {inst / frames:7.1f} warp-inst, "
                  f"{wf / frames:6.1f} L1 wavefronts ({wfs / frames:5.1f} shared)")


if __name__ == "__main__":
    main()
