"""gpurun_out/<tag>_traffic_<name>.csv (tools/ncu_traffic.sh) -> profiles/traffic.json + the raw CSVs under profiles/.

    python tools/ncu_traffic_json.py <tag>
"""
import csv
import json
import os
import shutil
import sys

tag = sys.argv[1]
NAMES = {"k_classify_f32": "classify_f32", "k_pass_z": "edt_pass_z", "k_pass_x": "edt_pass_x", "k_pass_y": "edt_pass_y",
         "k_cell_measures": "cell_measures"}
WORKLOADS = {"twist512": "twist512x512x512", "assembly1024": "assembly1024x1024x1024",
             "assembly1024_slab8": "assembly1024x1024x1024/slab8"}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "usecond": 1e-6,
         "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
caps = []
for name, wl in WORKLOADS.items():
    path = f"gpurun_out/{tag}_traffic_{name}.csv"
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kernels = {}
    for r in rows[2:]:
        kn = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").strip()
        if kn not in NAMES:
            continue

        def val(m):
            return float(r[col[m]].replace(",", "")) * SCALE.get(units[col[m]], 1.0)
        kernels[NAMES[kn]] = {"dram_read_gb": val("dram__bytes_read.sum") / 1e9, "dram_write_gb": val("dram__bytes_write.sum") / 1e9,
                              "ncu_ms": val("gpu__time_duration.sum") * 1e3,
                              "sm_throughput_pct": float(r[col["sm__throughput.avg.pct_of_peak_sustained_elapsed"]]),
                              "warp_instructions": val("smsp__inst_executed.sum")}
    dst = f"profiles/{tag}_traffic_{name}.csv"
    shutil.copy(path, dst)
    caps.append({"workload": wl, "source": f"{dst} (ncu --set full --clock-control none, one whole-{'slab' if 'slab' in name else 'grid'} "
                 "launch per kernel, VC_WORKERS=0, one B200)", "kernels": kernels})
json.dump({"captures": caps}, open("profiles/traffic.json", "w"), indent=1)
for c in caps:
    print(c["workload"], {k: round(v["dram_read_gb"] + v["dram_write_gb"], 3) for k, v in c["kernels"].items()})
