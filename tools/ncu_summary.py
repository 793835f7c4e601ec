"""Summarise an ncu report (read here, on the CPU box): python tools/ncu_summary.py <report.ncu-rep> <out.txt> [traffic.json]

Writes the metrics the roofline discussion uses (with their units) for every captured launch and,
optionally, the DRAM bytes per launch of the first one as JSON for bench.py's roofline.traffic."""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_config_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    lines, traffic = [], None
    for r in launches:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"== {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:75s} {d[k]:>18s} {u[k]}")
        for k in sorted(d):
            if "issue_stalled" in k and k.endswith("per_warp_active.pct"):
                try:
                    if float(d[k]) >= 2.0:
                        lines.append(f"{k:75s} {d[k]:>18s} {u[k]}")
                except ValueError:
                    pass
        if traffic is None:
            rd = float(d["dram__bytes_read.sum"]) * SCALE[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * SCALE[u["dram__bytes_write.sum"]]
            traffic = {"kernel": d.get("Kernel Name"), "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "dram_bytes_per_launch": rd + wr, "duration_ns_under_ncu": float(d["gpu__time_duration.sum"]) *
                       {"ns": 1, "us": 1e3, "ms": 1e6}.get(u["gpu__time_duration.sum"], 1), "report": rep}
    open(out, "w").write("\n".join(lines) + "\n")
    if len(sys.argv) > 3 and traffic:
        json.dump(traffic, open(sys.argv[3], "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
