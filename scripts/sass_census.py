"""SASS census of the shipped library: per kernel, how many tcgen05 / TMEM / TMA / legacy-MMA instructions it holds
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG/UBLKCP, mma.sync -> HMMA).

    python scripts/sass_census.py [out.txt]        (runs cuobjdump -sass on gram_b200/lib/libgram_b200.so; no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gram_b200", "lib", "libgram_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "HMMA", "LDGSTS", "SYNCS"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_census.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["_total"] += 1
            if op in MNEMONICS:
                counts[cur][op] += 1
    lines = [f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a); instruction counts per kernel; blank = 0",
             "# kernel".ljust(100) + "".join(m.rjust(9) for m in MNEMONICS) + "    total"]
    tot = collections.Counter()
    for fn in order:
        c = counts[fn]
        name = re.sub(r"\s+", " ", demangle(fn))
        name = re.sub(r"\((int|bool)\)", "", name)             # template-argument casts
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("gram::", "")[:98]
        lines.append(name.ljust(100) + "".join((str(c[m]) if c[m] else "").rjust(9) for m in MNEMONICS) + str(c["_total"]).rjust(9))
        tot.update(c)
    lines.append("ALL KERNELS".ljust(100) + "".join(str(tot[m]).rjust(9) for m in MNEMONICS) + str(tot["_total"]).rjust(9))
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[-1:]), "->", out)


if __name__ == "__main__":
    main()
