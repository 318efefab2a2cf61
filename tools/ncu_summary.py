"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/<tag>_launches.csv  > profiles/<tag>_launches.txt
    python tools/ncu_summary.py full     gpurun_out/<tag>_prof.ncu-rep  > profiles/<tag>_ncu_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        v = float(r[mi].replace(",", ""))
        u = r[ui]
        v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v * 1e3 if u in ("s", "second") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)-1} launches, {tot:.3f} ms of kernel time (ncu: cold-cache, serialised; compare SHARES)")
    print(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'ms/launch':>10s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:58s} {v[0]:8d} {v[1]:10.3f} {v[1]/v[0]:10.4f} {v[1]/tot:6.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        print(f"== {name}")
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"   {m:85s} {r[i]:>16s} {units[i]}")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            u = units[hdr.index("dram__bytes_read.sum")]
            t = float(r[hdr.index("gpu__time_duration.sum")]); tu = units[hdr.index("gpu__time_duration.sum")]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
            ts = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}[tu.replace("second", "s") if "second" in tu else tu]
            print(f"   traffic (dram read+write) = {(rd+wr)*scale/1e9:.4f} GB  ->  {(rd+wr)*scale/1e9/(t*ts):.1f} GB/s under ncu")
        except Exception as ex:
            print("   traffic: n/a", ex)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
