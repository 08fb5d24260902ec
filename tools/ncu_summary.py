"""Print the key metrics of every kernel in an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__inst_executed_op_global_ld.sum"]
STALL = "smsp__average_warps_issue_stalled_"
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(h, r))
    print("=" * 100)
    print(d.get("Kernel Name", "?")[:120])
    for k in KEYS:
        if k in d:
            print(f"  {k:62s} {d[k]:>16s} {units[h.index(k)]}")
    st = sorted(((float(v.replace(',', '')), k[len(STALL):-len('_per_issue_active.ratio')]) for k, v in d.items()
                 if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
    print("  stalls (warps per issue-active cycle):", ", ".join(f"{n} {v:.2f}" for v, n in st[:7]))
