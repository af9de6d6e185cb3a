"""Turn gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and gpurun_out/prof_volume.ncu-rep (ncu --set full)
into the tracked summaries under profiles/ (per round) and profiles/traffic.json (read by bench.py for roofline.traffic).
    python scripts/summarize_profiles.py r01"""
import csv, json, os, re, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list -> per-kernel totals and shares
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", "launches.csv"))) if len(r) > 10]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"cb200::<unnamed>::|void ", "", name)
    t = float(r[iv].replace(",", ""))
    tot[name][0] += 1
    tot[name][1] += t
unit = [r for r in rows[1:] if r[im] == "gpu__time_duration.sum"][0][hdr.index("Metric Unit")]
total = sum(v[1] for v in tot.values())
with open(os.path.join(out, "%s_launches_summary.txt" % tag), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 3 --warmup 3 --no-extras\n")
    f.write("(cold-cache, serialised launches: compare SHARES, not absolutes)  unit: %s\n\n" % unit)
    f.write("%-62s %8s %14s %8s\n" % ("kernel", "launches", "total", "share"))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write("%-62s %8d %14.1f %7.1f%%\n" % (k[:62], v[0], v[1], 100 * v[1] / total))
print(open(os.path.join(out, "%s_launches_summary.txt" % tag)).read())

# ---- full capture -> key metrics per captured launch
rr = []
for rep in ("prof_volume.ncu-rep", "prof_ops.ncu-rep"):
    path = os.path.join(ROOT, "gpurun_out", rep)
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    rr = part if not rr else rr + part[2:]
h, u = rr[0], rr[1]
idx = {k: i for i, k in enumerate(h)}
want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
traffic = {}
with open(os.path.join(out, "%s_volume_kernel_ncu.txt" % tag), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:volume_action_ws: -s 4 -c 1 python bench.py --steps 3 --warmup 3 --no-extras (fused kernel);\n"
            "-s 2 -c 2 python scripts/prof_ops.py 1024 5 (stand-alone stiffness / weighted mass)\n\n")
    for r in rr[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void cb200::<unnamed>::", "")
        f.write("== %s\n" % name)
        for k in want:
            if k in idx:
                f.write("   %-68s %s %s\n" % (k, r[idx[k]], u[idx[k]]))
        def gb(k):
            v, un = float(r[idx[k]].replace(",", "")), u[idx[k]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[un]
        m = re.search(r"volume_action_(?:kernel|ws)<(?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(bool\))?(\d)(?:, (?:\(int\))?(\d+))?(?:, (?:\(int\))?\d+)?>", r[idx["Kernel Name"]])
        if m:
            if m.group(4) and m.group(4) != "0":
                key = "helmholtz_%s_%s_%s_nx1024" % (m.group(1), m.group(2), m.group(4))
            else:
                key = ("stiffness" if m.group(3) == "1" else "mass") + "_%s_%s_nx1024" % (m.group(1), m.group(2))
            traffic[key] = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
json.dump(traffic, open(os.path.join(out, "traffic.json"), "w"), indent=1)
print(open(os.path.join(out, "%s_volume_kernel_ncu.txt" % tag)).read()[:3000])
print(traffic)
