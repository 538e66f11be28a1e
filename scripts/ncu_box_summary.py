"""Runs ON THE GPU BOX at the end of a profiling call: turns gpurun_out/*.ncu-rep into small markdown summaries (raw ncu
metrics per kernel launch) next to them and deletes the reports, so that gpurun_out/ stays under the 64 MiB that are copied
back.  usage: python scripts/ncu_box_summary.py <tag>      (then copy gpurun_out/<tag>_*.md into profiles/)"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__clock_rate? "]


def short(name):
    return re.sub(r"\(.*", "", name.replace("void ", "").replace("tg::", ""))


for rep in sorted(glob.glob(os.path.join(GO, f"{tag}_prof_*.ncu-rep"))):
    title = os.path.basename(rep)[len(tag) + 6:-8]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(GO, f"{tag}_{title}.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none ({tag}, {os.path.basename(rep)})\n\n")
        seen = collections.Counter()
        for r in rows[2:]:
            name = short(r[idx["Kernel Name"]])
            seen[name] += 1
            if seen[name] > 2:
                continue
            f.write(f"## {name} (launch {seen[name]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            try:
                rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")); wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
                f.write(f"| dram traffic (read+write) | {rd + wr:.3f} | {units[idx['dram__bytes_read.sum']]} |\n")
            except Exception:
                pass
            f.write("\n")
    os.remove(rep)
for src in sorted(glob.glob(os.path.join(GO, f"{tag}_launches*.csv"))):
    rows = list(csv.reader(open(src, errors="ignore")))
    hdr, agg, total = None, collections.OrderedDict(), 0.0
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        ms = v / 1e6 if d["Metric Unit"].startswith("n") else (v / 1e3 if d["Metric Unit"].startswith("u") else v)
        a = agg.setdefault(short(d["Kernel Name"]), [0, 0.0]); a[0] += 1; a[1] += ms; total += ms
    with open(src[:-4] + ".md", "w") as f:
        f.write(f"# ncu launch list ({os.path.basename(src)}): `ncu --metrics gpu__time_duration.sum --clock-control none ...`\n\n")
        f.write("Per-launch times are cold-cache and serialised; compare SHARES with bench.py's `roofline.kernel_share_of_step`.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {n} | {ms:.3f} | {ms / total * 100:.2f}% |\n")
        f.write(f"| total | | {total:.3f} | |\n")
    if os.path.getsize(src) > (1 << 20):
        os.remove(src)
print(sorted(os.listdir(GO)))
