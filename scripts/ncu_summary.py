"""Summarise ncu captures from gpurun_out/ into profiles/ (tracked).  Usage: python scripts/ncu_summary.py <round-tag>"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(OUT, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def short(name):
    name = name.replace("void ", "").replace("tg::", "")
    return re.sub(r"\(.*", "", name)


def launches():
    src = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if not os.path.isfile(src):
        return
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
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --also-19 0 --also-dedup 0`\n\n")
        f.write("Per-launch times are cold-cache and serialised; compare SHARES with bench.py's `roofline.kernel_share_of_step`.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {n} | {ms:.3f} | {ms / total * 100:.2f}% |\n")
        f.write(f"| total | | {total:.3f} | |\n")


def full(rep, title):
    src = os.path.join(ROOT, "gpurun_out", rep)
    if not os.path.isfile(src):
        return
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(OUT, f"{tag}_{title}.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none ({tag}, {rep})\n\n")
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
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")); wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            f.write(f"| dram traffic (read+write) | {rd + wr:.3f} | {units[idx['dram__bytes_read.sum']]} |\n\n")


launches()
full("prof_dualnet.ncu-rep", "dualnet_tc")
full("prof_search.ncu-rep", "search_kernels")
print(os.listdir(OUT))
