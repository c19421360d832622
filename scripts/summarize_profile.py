"""Turn ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/summarize_profile.py <tag> <launches.csv[.gz]> [<report.ncu-rep | raw-page.csv> ...]
"""
import collections
import csv
import gzip
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(tag, path):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(io.StringIO("".join(lines)))
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    n = 0
    for row in r:
        v = float(row[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row[ui], 1e-6)
        a = agg.setdefault(row[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    out = os.path.join(ROOT, "profiles", f"{tag}_launches_summary.txt")
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; {n} launches, {tot:.3f} ms total\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'ms':>10s} {'share':>7s} {'count':>6s}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:10.3f} {100 * v[1] / tot:6.1f}% {v[0]:6d}  {k[:150]}\n")
    dst_path = os.path.join(ROOT, "profiles", f"{tag}_launches.csv.gz")
    if path.endswith(".gz"):
        shutil.copyfile(path, dst_path)
    else:
        with open(path, "rb") as src, gzip.open(dst_path, "wb") as dst:
            shutil.copyfileobj(src, dst)
    print("wrote", out)


def full(tag, reps):
    out = os.path.join(ROOT, "profiles", f"{tag}_full_metrics.txt")
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (one block per captured launch)\n")
        for rep in reps:
            full_one(f, rep)
    print("wrote", out)


STALLS = 6   # top warp-stall reasons listed per launch


def full_one(f, rep):
    if rep.endswith(".csv"):
        raw = open(rep).read()      # `ncu -i report --page raw --csv`, exported on the GPU box
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    if True:
        for r in rows[2:]:
            f.write(f"\n== {r[idx['Kernel Name']][:120]}\n")
            for m in KEEP:
                if m in idx:
                    f.write(f"  {m:72s} {r[idx[m]]} {units[idx[m]]}\n")
            stalls = []
            for h, i in idx.items():
                if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a"):
                    stalls.append((float(r[i].replace(",", "")), h.split("issue_stalled_")[1].split("_per_issue")[0]))
            f.write("  top stalls (warps per issue-active cycle): "
                    + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:STALLS]) + "\n")


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 3:
        full(sys.argv[1], sys.argv[3:])
