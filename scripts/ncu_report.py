"""Text summary of every launch in one or more .ncu-rep files (ncu --set full): key metrics (ncu --page raw) and, per launch, the
warp-stall samples of the helper and of the compute branch (ncu --page source, split at USETMAXREG.TRY_ALLOC).
    python scripts/ncu_report.py "<command line that was profiled>" a.ncu-rep [b.ncu-rep ...] > profiles/xxx.txt"""
import csv, io, re, subprocess, sys
cmd, reps = sys.argv[1], sys.argv[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(cmd)
print()
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    idx = {k: i for i, k in enumerate(h)}
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(src)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and row:
            cur["rows"].append(row)
    for li, r in enumerate(rr[2:]):
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]].replace("(int)", "").replace("(bool)", "")).replace("void cb200::<unnamed>::", "").replace("void unnamed>::", "")
        print("== %s   [%s]" % (name, rep.split("/")[-1]))
        for k in want + [k for k in h if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k]:
            if k in idx:
                v = r[idx[k]]
                try:
                    if "issue_stalled" in k and float(v) < 0.05:
                        continue
                except ValueError:
                    pass
                print("   %-80s %-14s %s" % (k, u[idx[k]], v))
        if li < len(blocks):
            b = blocks[li]
            hh = b["rows"][0]
            ix = {k: i for i, k in enumerate(hh)}
            rows = b["rows"][1:]
            cols = [k for k in hh if k.startswith("stall_") and "Not Issued" not in k]
            split = next((i for i, q in enumerate(rows) if "USETMAXREG.TRY_ALLOC" in q[ix["Source"]]), None)
            parts = (("helper branch", rows[:split]), ("compute branch", rows[split:])) if split is not None else (("all warps", rows),)
            for nm, rs in parts:
                ag = {k[6:]: sum(int(q[ix[k]]) for q in rs) for k in cols}
                tot = sum(ag.values())
                print("   warp-stall samples, %s: %d  " % (nm, tot) + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(ag.items(), key=lambda kv: -kv[1]) if v * 200 >= tot))
        print()
