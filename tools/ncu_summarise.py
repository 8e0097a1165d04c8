"""Summaries of an .ncu-rep for profiles/: details page as CSV + selected raw metrics.
usage: python tools/ncu_summarise.py gpurun_out/x.ncu-rep profiles/r1c_ncu_full_k3"""
import csv, subprocess, sys

SELECT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
          "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__average_warp_latency_per_inst_issued.ratio")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    open(out + "_details.csv", "w").write(det)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keep = [i for i, n in enumerate(hdr) if n == "Kernel Name" or any(n.startswith(s) for s in SELECT)]
    with open(out + "_raw_selected.csv", "w", newline="") as f:
        w = csv.writer(f)
        for r in (hdr, units, vals):
            w.writerow([r[i] for i in keep])
    for i in keep:
        print(hdr[i], units[i], vals[i])


if __name__ == "__main__":
    main()
