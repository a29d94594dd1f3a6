"""Static evidence from the built library (no GPU needed): per kernel the registers / shared memory / spills ptxas
reports (`cuobjdump -res-usage`) and the count of the SASS mnemonics that show which hardware path it uses
(UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier).

    python scripts/sass_check.py > profiles/r02/sass_resources.txt
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mp_reid_b200", "libmpreid_b200.so")
MNEMONICS = ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "MATCH", "REDUX", "ATOMG", "RED", "ATOMS", "FFMA", "DFMA", "HFMA2", "SHFL", "BAR")


def demangle(names):
    out = subprocess.run(["c++filt", "-p"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage, fn = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            usage[fn] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL)\[?\d*\]?:(\d+)", line)}
            fn = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    pat = re.compile(r"\b(" + "|".join(MNEMONICS) + r")\b")
    counts, fn = collections.defaultdict(collections.Counter), None
    n_instr = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            n_instr[fn] += 1
            mm = pat.search(line)
            if mm:
                counts[fn][mm.group(1)] += 1
    names = demangle(sorted(usage))
    # the default instantiations only for the heavily templated GEMM: 3xFP16 (PREC 3), sq-euclid (METRIC 0)
    def keep(d):
        return "k_dist_tc<" not in d or re.search(r"k_dist_tc<3, 128, 0, (true|false), (true|false), (true|false)>", d)
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(usage)} kernels; k_dist_tc template: <PREC, ROW_BYTES, METRIC, VEC, CTA2, FUSE>, of its "
          f"{sum('k_dist_tc<' in d for d in names.values())} instantiations only 3xFP16 / sq-euclid are listed")
    print(f"{'kernel':100s} {'regs':>4s} {'stack':>5s} {'smem':>6s} {'instr':>6s}  mnemonics")
    for fn in sorted(usage, key=lambda f: names[f]):
        d = names[fn]
        if not keep(d):
            continue
        u = usage[fn]
        d = re.sub(r"\(.*$", "", d)
        print(f"{d[:100]:100s} {u.get('REG', 0):4d} {u.get('STACK', 0):5d} {u.get('SHARED', 0):6d} {n_instr[fn]:6d}  "
              + " ".join(f"{k}={v}" for k, v in sorted(counts[fn].items())))
    spills = [names[f] for f, u in usage.items() if u.get("STACK", 0) > 0]
    print(f"# kernels with a stack frame (spills or local arrays): {len(spills)}")
    for s in sorted(spills):
        print("#   " + re.sub(r"\(.*$", "", s)[:120])


if __name__ == "__main__":
    sys.exit(main())
