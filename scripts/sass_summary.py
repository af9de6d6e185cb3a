"""SASS evidence for profiles/: opcode histogram + the lines that prove the hardware paths (UBLKCP = cp.async.bulk / TMA bulk copy,
SYNCS = mbarrier, USETMAXREG = setmaxnreg, UCGABAR = cluster barrier, DFMA / FFMA, LDCU = uniform constant loads, SHFL) of the
named kernels, from `cuobjdump -sass` of the built library. No GPU needed.
    python scripts/sass_summary.py <substring of the mangled kernel name> [...]  > profiles/rNN_sass_<kernel>.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cuddhelmholtz_b200", "lib", "libcuddh_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for pat in sys.argv[1:]:
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if pat not in name:
            continue
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        ops = collections.Counter()
        lines = []
        for ln in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                ops[m.group(3).split(".")[0]] += 1
                lines.append(ln.rstrip())
        total = sum(ops.values())
        print("== %s\n   (%s)\n   %d SASS instructions (cuobjdump -sass, sm_100a)" % (demangled[:200], name[:120], total))
        print("   " + ", ".join("%s %d" % kv for kv in ops.most_common(28)))
        for key in ("USETMAXREG", "UBLKCP", "UBLKPF", "SYNCS", "UCGABAR", "FENCE", "LDGSTS", "RED", "SHFL", "BAR"):
            hits = [l for l in lines if re.search(r"\b" + key, l)]
            if hits:
                print("   -- %s x%d, e.g." % (key, len(hits)))
                for l in hits[:3]:
                    print("      " + re.sub(r"\s+", " ", l.strip())[:150])
        fma = [l for l in lines if re.search(r"\b(DFMA|FFMA)\b", l)]
        if fma:
            print("   -- first FMA lines:")
            for l in fma[:6]:
                print("      " + re.sub(r"\s+", " ", l.strip())[:150])
        print()
