"""Digest of an ncu report brought back in gpurun_out/: headline metrics, warp-stall breakdown and the hottest SASS lines.

    python profiles/ncu_digest.py gpurun_out/r2_gn64occ2.ncu-rep [n_hot]
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        print("kernel:", d[hdr.index("Kernel Name")][:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:75s} {d[i]} {units[i]}")


def source(path, n_hot):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda v: int(v) if v.isdigit() else 0
    total = sum(num(d[idx["# Samples"]]) for d in data)
    tot = {s: sum(num(d[idx[s]]) for d in data) for s in stalls}
    print(f"  warp-stall samples: {total}")
    print("  " + ", ".join(f"{s[6:]} {100 * v / max(total, 1):.1f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
    for d in sorted(data, key=lambda d: -num(d[idx["# Samples"]]))[:n_hot]:
        top = sorted(((num(d[idx[s]]), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"  {d[idx['# Samples']]:>6} {d[idx['Instructions Executed']]:>9}  {d[idx['Source']].strip()[:72]:72s} {top}")


if __name__ == "__main__":
    raw(sys.argv[1])
    source(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
