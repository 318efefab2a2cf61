"""Hot SASS lines of a kernel from an ncu report's source page.
    python tools/ncu_hot.py <rep> <kernel-regex> [top=40] [launch-skip=0]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break  # next kernel's section
    data.append(r)
si, src, ie, at = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in data); toti = sum(int(r[ie]) for r in data)
print(rows[0][1][:100]); print("samples", tot, "warp-instr", toti, "sass lines", len(data))
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stalls}
print("stall mix:", {k: f"{100*v/max(tot,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]})
order = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:top]
for i in sorted(order):
    r = data[i]
    st = max(stalls, key=lambda k: int(r[k]))
    print(f"{i:5d} {100*int(r[si])/tot:5.1f}% exec={int(r[ie]):9d} thr={r[at]:>5s} {hdr[st]:>14s} | {r[src].strip()[:100]}")
