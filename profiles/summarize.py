"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launch_shares.md
    python profiles/summarize.py full     gpurun_out/prof.ncu-rep     > profiles/r1_conv_tc_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<.*", "", r["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r["Metric Unit"], 1.0)
        agg.setdefault(name, []).append(v)
    total = sum(sum(v) for v in agg.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none, {len(rows)} launches, {total:.3f} ms total "
          "(cold-cache, serialised: compare SHARES)\n")
    print("| kernel | launches | total ms | share | max ms |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k[:70]}` | {len(v)} | {sum(v):.3f} | {100 * sum(v) / total:.1f}% | {max(v):.3f} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    name_i = hdr.index("Kernel Name")
    print("ncu --set full --clock-control none (one row per captured launch)\n")
    print("| # | " + " | ".join(w for w, _ in idx) + " |\n|---|" + "---|" * len(idx))
    for n, d in enumerate(data):
        print(f"| {n} | " + " | ".join(f"{d[i]} {units[i]}".strip() for _, i in idx) + " |")
    print("\nkernel:", re.sub(r"\(.*", "", data[0][name_i]) if data else "-")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
