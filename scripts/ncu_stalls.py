"""Where a kernel's warps wait: the sampled stall PCs of an `ncu --set full --import-source on` capture, grouped by SASS opcode
and listed with their neighbourhood (the instruction a warp is sampled AT is the one that cannot issue; what it waits for is
usually one or two instructions above it).

    python scripts/ncu_stalls.py gpurun_out/prof_x.ncu-rep [kernel-index] [top-n]
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    heads = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
    if not heads:
        print("no source page in", rep)
        return
    print(f"{len(heads)} kernels in the report; showing #{which}:", rows[heads[which] - 1][1][:90] if heads[which] else "")
    h = rows[heads[which]]
    end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
    si, ci = h.index("Source"), h.index("# Samples")
    body = []
    for r in rows[heads[which] + 1:end]:
        try:
            body.append((float(r[ci]), r[si].strip()))
        except (ValueError, IndexError):
            pass
    tot = sum(v for v, _ in body) or 1.0
    by_op = collections.Counter()
    for v, src in body:
        m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
        by_op[m.group(1) if m else "?"] += v
    print(f"{int(tot)} samples over {len(body)} instructions")
    print("by opcode:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in by_op.most_common(12)))
    order = sorted(range(len(body)), key=lambda i: -body[i][0])[:top]
    for i in order:
        print(f"\n--- {100 * body[i][0] / tot:.1f}% at instruction {i}")
        for j in range(max(0, i - 4), min(len(body), i + 2)):
            print(f"   {'>>' if j == i else '  '} {100 * body[j][0] / tot:5.1f}%  {body[j][1][:110]}")


if __name__ == "__main__":
    main()
