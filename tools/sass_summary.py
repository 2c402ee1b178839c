"""Evidence that libsce.so is tcgen05 / TMEM / TMA code, per kernel (run in the build container, no GPU needed):
  python tools/sass_summary.py > profiles/r02_sass_summary.txt
Part 1: counts of the SASS mnemonics of B200_PROFILING.md per kernel (cuobjdump -sass of the in-tree library).
Part 2: registers / spills / shared memory per kernel from cuobjdump -res-usage (what `ptxas -v` prints at build time)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sparse_coding_b200", "libsce.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "HMMA", "SYNCS"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(sce::GemmParams<.*", "", name)
    name = name.replace("void sce::", "").replace("sce::", "")
    return name[:110]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
    names = demangle(list(counts))
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a)")
    print("# UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4, UTMALDG/UTMASTG = TMA load/store, LDTM = tcgen05.ld,")
    print("# UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0)")
    print(f"{'kernel':112s} {'instrs':>7s} " + " ".join(f"{m:>8s}" for m in MNEMONICS))
    tot = collections.Counter()
    for k, c in counts.items():
        if c["_total"] == 0:
            continue
        print(f"{short(names[k]):112s} {c['_total']:7d} " + " ".join(f"{c[m]:8d}" for m in MNEMONICS))
        tot.update(c)
    print(f"{'TOTAL':112s} {tot['_total']:7d} " + " ".join(f"{tot[m]:8d}" for m in MNEMONICS))
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    print("\n# resource usage per kernel (cuobjdump -res-usage == ptxas -v): registers, spill stack, static shared memory")
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            reg = re.search(r"REG:(\d+)", line).group(1)
            stack = re.search(r"STACK:(\d+)", line).group(1)
            shared = re.search(r"SHARED:(\d+)", line).group(1)
            local = re.search(r"LOCAL:(\d+)", line).group(1)
            nm = demangle([fn])[fn]
            print(f"{short(nm):112s} REG {reg:>4s}  STACK {stack:>5s}  LOCAL {local:>4s}  SHARED(static) {shared:>6s}")
            fn = None


if __name__ == "__main__":
    main()
