"""compact key metrics of every kernel in an .ncu-rep:  python tools/ncu_keys.py rep.ncu-rep"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__inst_executed_op_global_ld.sum"]
STALL = "smsp__average_warps_issue_stalled_"
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
want = sys.argv[2] if len(sys.argv) > 2 else None
for r in rows[2:]:
    d = dict(zip(h, r))
    print("=" * 100)
    print(d.get("Kernel Name", "?")[:120])
    for k in h:
        if k in KEYS or (want and want in k):
            print(f"  {k:66s} {d[k]:>16s} {units[h.index(k)]}")
    st = sorted(((float(v.replace(',', '')), k[len(STALL):-len('_per_issue_active.ratio')]) for k, v in d.items()
                 if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
    print("  stalls (warps per issue-active cycle):", ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))
