"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag>          e.g. r1

Reads  gpurun_out/launches_<tag>.csv                 (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/prof_*_<tag>.ncu-rep               (ncu --set full captures)
Writes profiles/<tag>_launches.csv                   (the launch list itself, trimmed columns)
       profiles/<tag>_launch_summary.txt             (per-kernel totals and SHARES of the step)
       profiles/<tag>_<capture>_metrics.csv          (selected raw metrics per captured launch)
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__cluster_size",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
]


def launches(tag):
    path = os.path.join(SRC, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    # keep exactly one step: from the last passage_len_kernel (first launch of a generate call) that is followed by a
    # beam_finalize_kernel (its last launch) through that launch
    starts = [i for i, r in enumerate(data) if "passage_len_kernel" in r[ki]]
    ends = [i for i, r in enumerate(data) if "beam_finalize_kernel" in r[ki]]
    whole = [(a, min(e for e in ends if e > a)) for a in starts if any(e > a for e in ends)]
    if whole:
        a, e = whole[-1]
        data = data[a:e + 1]
    with open(os.path.join(OUT, f"{tag}_launches.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration_ns"])
        for r in data:
            w.writerow([r[0], re.sub(r"\(.*", "", r[ki]).replace("void ", ""), r[gi], r[bi], r[vi]])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(r[vi]) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launch_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; {len(data)} consecutive launches of bench.py = one whole step (generate call)\n")
        f.write("# cold-cache, serialised per-launch times: compare SHARES with bench.py's kernel_classes, not absolutes\n")
        f.write(f"# total {tot:.1f} us\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d} avg={v[1] / v[0]:8.1f} us  {k}\n")
    print(open(os.path.join(OUT, f"{tag}_launch_summary.txt")).read())


def captures(tag):
    for rep in sorted(glob.glob(os.path.join(SRC, f"prof_*_{tag}.ncu-rep"))):
        name = os.path.basename(rep)[len("prof_"):-len(f"_{tag}.ncu-rep")]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        cols = [("Kernel Name", hdr.index("Kernel Name"))] + [(m, hdr.index(m)) for m in METRICS if m in hdr]
        path = os.path.join(OUT, f"{tag}_{name}_metrics.csv")
        with open(path, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow([c for c, _ in cols])
            w.writerow([units[i] for _, i in cols])
            for r in rows[2:]:
                w.writerow([re.sub(r"\(.*", "", r[i]) if c == "Kernel Name" else r[i] for c, i in cols])
        print("wrote", path)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    captures(tag)
