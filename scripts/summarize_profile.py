"""Turn ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/summarize_profile.py <tag> <launches.csv> [<report.ncu-rep>]
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
    with open(path) as f:
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
    with open(path, "rb") as src, gzip.open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv.gz"), "wb") as dst:
        shutil.copyfileobj(src, dst)
    print("wrote", out)


def full(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = os.path.join(ROOT, "profiles", f"{tag}_full_metrics.txt")
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (one block per captured launch)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[idx['Kernel Name']][:120]}\n")
            for m in KEEP:
                if m in idx:
                    f.write(f"  {m:72s} {r[idx[m]]} {units[idx[m]]}\n")
    print("wrote", out)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 3:
        full(sys.argv[1], sys.argv[3])
