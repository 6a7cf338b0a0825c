#!/usr/bin/env python
"""Compact per-launch summary of an .ncu-rep (run here, no GPU needed):
   ncu_summary.py report.ncu-rep out.csv ["header comment"] [traffic.json key]
Writes one column per profiled launch with the metrics DESIGN.md / bench.py cite; with a 4th/5th argument
also records dram read+write bytes per launch under `key` in the JSON file bench.py reads for
roofline.traffic."""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    comment = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = ["%s#%s" % (r[idx["Kernel Name"]].split("(")[0][:40], r[idx["ID"]]) for r in data]
    with open(out, "w") as f:
        if comment:
            f.write("# " + comment + "\n")
        f.write("metric,unit," + ",".join(names) + "\n")
        for m in METRICS:
            if m in idx:
                f.write("%s,%s,%s\n" % (m, units[idx[m]], ",".join(r[idx[m]].replace(",", "") for r in data)))
    if len(sys.argv) > 5 and sys.argv[5]:
        path, key = sys.argv[4], sys.argv[5]
        try:
            js = json.load(open(path))
        except Exception:
            js = {}
        ent = []
        for r in data:
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = r[idx[m]].replace(",", "")
                tot += float(v) * UNIT.get(units[idx[m]], 1.0) if v not in ("", "nan", "-nan") else float("nan")
            ent.append({"kernel": r[idx["Kernel Name"]].split("(")[0], "dram_bytes": tot,
                        "ms": float(r[idx["gpu__time_duration.sum"]].replace(",", "")) *
                        {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0)})
        js[key] = {"launches": ent, "source": out.split("/")[-1]}
        json.dump(js, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
