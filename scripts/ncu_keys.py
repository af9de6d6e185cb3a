"""Print the key metrics of every launch in an .ncu-rep (ncu --page raw --csv), one column per launch."""
import csv, subprocess, sys, re
rep = sys.argv[1]
extra = sys.argv[2:] 
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u = rr[0], rr[1]
idx = {k: i for i, k in enumerate(h)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "idc__requests.sum", "idc__request_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
want += [k for k in h if any(re.search(e, k) for e in extra)]
for k in want:
    if k in idx:
        print("%-84s %-8s %s" % (k, u[idx[k]], " | ".join(r[idx[k]] for r in rr[2:])))
print("kernels:", " | ".join(re.sub(r"\(.*", "", r[idx["Kernel Name"]])[-60:] for r in rr[2:]))
