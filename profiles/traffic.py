"""Per-kernel-class time and DRAM traffic of ONE composed evaluation, from an ncu launch list.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -c 400 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline
    python profiles/traffic.py gpurun_out/traffic.csv > profiles/r1_evaluation_traffic.json
    python profiles/traffic.py gpurun_out/traffic.csv --md > profiles/r1_evaluation_launches.md

One evaluation = the launches from one stem_kernel up to (not including) the next one.
"""
import collections
import csv
import json
import re
import sys


def main(path):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(r["Metric Unit"], 1.0)
        d[r["Metric Name"]] = v * scale
    seq = [launches[k] for k in sorted(launches)]
    stems = [i for i, d in enumerate(seq) if "stem_kernel" in d["name"] or "stem_mma_kernel" in d["name"]]
    if len(stems) < 2:
        raise SystemExit("need at least two stem launches in the capture")
    ev = seq[stems[0]:stems[1]]
    classes = collections.OrderedDict()
    for d in ev:
        name = re.sub(r"<.*", "", d["name"]).replace("void ", "").replace("cindm::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        c = classes.setdefault(name, {"launches": 0, "ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        c["launches"] += 1
        c["ms"] += d.get("gpu__time_duration.sum", 0.0)
        c["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
        c["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
    total = sum(c["ms"] for c in classes.values())
    for c in classes.values():
        c["share"] = c["ms"] / total
    out = {"evaluation_kernels": len(ev), "total_ms": total,
           "classes": collections.OrderedDict(sorted(classes.items(), key=lambda kv: -kv[1]["ms"]))}
    if "--md" in sys.argv:
        print("One composed-epsilon evaluation (C4 per GPU: S = 43 008 slices) between two consecutive stem launches.")
        print("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
              "(serialised, cold-cache: compare SHARES).\n")
        print("| kernel | launches | total ms | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|")
        for k, c in out["classes"].items():
            print(f"| `{k[:60]}` | {c['launches']} | {c['ms']:.3f} | {100 * c['share']:.1f}% | "
                  f"{c['dram_read_bytes'] / 1e6:.0f} | {c['dram_write_bytes'] / 1e6:.0f} |")
        print(f"\ntotal {out['total_ms']:.3f} ms over {out['evaluation_kernels']} kernels")
        return
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1])
