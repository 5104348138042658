"""Hottest SASS instructions of an ncu report (source page): samples, stall reasons and the instructions before them.
usage: ncu_hot.py report.ncu-rep [top] [kernels]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 8; nk = int(sys.argv[3]) if len(sys.argv) > 3 else 2
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for w in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"):
    if w in hdr:
        i = hdr.index(w); print(w, [r[i][:26] for r in rows[2:]])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = []; cur = None; hdr = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "data": []}; blocks.append(cur); continue
    if r and r[0] == "Address": hdr = r; continue
    if cur is not None and hdr and len(r) >= 10: cur["data"].append(r)
si = hdr.index("# Samples"); src = hdr.index("Source")
cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
ci = [hdr.index(c) for c in cols]
for b in blocks[:nk]:
    d = b["data"]; tot = {c: 0 for c in cols}; n = 0
    for r in d:
        n += int(r[si] or 0)
        for c, i in zip(cols, ci): tot[c] += int(r[i] or 0)
    print("=====", b["name"][:40], "samples", n, {k[6:]: v for k, v in tot.items() if v > n * 0.02})
    for k in sorted(range(len(d)), key=lambda k: -int(d[k][si] or 0))[:top]:
        r = d[k]
        print("--", r[si], r[src][:80], {c[6:]: int(r[i]) for c, i in zip(cols, ci) if int(r[i] or 0)})
        for kk in range(max(0, k - 3), k): print("        ", d[kk][src][:90])
