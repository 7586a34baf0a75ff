"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics quoted in profiles/."""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("kernel:", d["Kernel Name"].split("(")[0])
    for w in WANT:
        if w in d and d[w] not in ("", "n/a"):
            print(f"  {w} [{u[w]}] = {d[w]}")
    st = sorted([(float(d[h].replace(",", "")), h) for h in hdr if "average_warps_issue_stalled" in h and "not_issued" not in h and d[h] not in ("", "n/a")], reverse=True)[:6]
    for v, h in st:
        print(f"  stall {h.split('stalled_')[1].split('_per')[0]} = {v:.3f} warps per issue")
