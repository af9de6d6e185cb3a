"""Per-instruction warp-stall samples of one kernel from an .ncu-rep (ncu --page source --csv): phase totals + top lines.
   python scripts/ncu_src.py rep.ncu-rep <kernel substring> [top N]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None and row:
        cur["rows"].append(row)
for b in blocks:
    if pat not in b["name"]:
        continue
    h = b["rows"][0]
    ix = {k: i for i, k in enumerate(h)}
    rows = b["rows"][1:]
    tot = sum(int(r[ix["# Samples"]]) for r in rows)
    print(b["name"][:100], "samples", tot, "instrs", len(rows))
    stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[ix[k]]) for r in rows) for k in stall_cols}
    print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    # cumulative samples by instruction index (to split phases), markers at BAR
    # the compute warpgroup starts at USETMAXREG.TRY_ALLOC: stall breakdown of helper / compute code separately
    split = next((i for i, r in enumerate(rows) if "USETMAXREG.TRY_ALLOC" in r[ix["Source"]]), None)
    if split is not None:
        for nm, rs in (("helper", rows[:split]), ("compute", rows[split:])):
            ag = {k[6:]: sum(int(r[ix[k]]) for r in rs) for k in stall_cols}
            print(nm, sum(ag.values()), {k: v for k, v in sorted(ag.items(), key=lambda kv: -kv[1]) if v})
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][ix["# Samples"]]))[:top]
    for i in sorted(order):
        r = rows[i]
        st = {k[6:]: int(r[ix[k]]) for k in stall_cols if int(r[ix[k]])}
        print("%4d %6s %-60s %s" % (i, r[ix["# Samples"]], r[ix["Source"]].strip()[:60], st))
    break
