"""Extract the metrics profiles/README.md quotes from an .ncu-rep into a small CSV.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/rNN_ncu_x_summary.csv
"""
import csv
import subprocess
import sys

KEEP = (
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            w.writerow(["kernel", "", d.get("Kernel Name", "")])
            w.writerow(["metric", "unit", "value"])
            for h, u in zip(hdr, units):
                if h in KEEP:
                    w.writerow([h, u, d[h]])
                elif h.startswith("smsp__average_warp_issue_stalled_") and h.endswith("_per_warp_active.pct"):
                    try:
                        if float(d[h].replace(",", "")) >= 2.0:
                            w.writerow([h, u, d[h]])
                    except ValueError:
                        pass


if __name__ == "__main__":
    main()
