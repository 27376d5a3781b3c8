"""Summarise one kernel of an .ncu-rep into a small JSON for profiles/ (the raw reports stay in gpurun_out/).
usage: python tools/ncu_summary.py <report.ncu-rep> <out.json> <chains> "<title>"
"""
import csv, io, json, subprocess, sys

KEEP = [
    "gpu__time_duration.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main():
    rep, out, chains, title = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, val = rows[0], rows[1], rows[2]
    d = {"_chains": chains, "_kernel": val[hdr.index("Kernel Name")], "_title": title, "_source": rep.split("/")[-1]}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            try:
                v = float(val[i].replace(",", ""))
            except ValueError:
                v = val[i]
            d[k] = {"unit": units[i], "value": v}
    json.dump(d, open(out, "w"), indent=1)
    print("wrote", out, len(d) - 4, "metrics")


if __name__ == "__main__":
    main()
