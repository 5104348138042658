"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals and shares."""
import collections, csv, sys
f = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
d = collections.defaultdict(list)
for r in rows:
    d[r[ki].split("(")[0].replace("void ", "")[:60] + " " + r[gi].replace(" ", "")].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in d.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us total ({tot / 1e3 / steps:.1f} us per step over {steps:g} steps)")
print("| kernel grid | launches | total us | share | avg us |")
print("|---|---|---|---|---|")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    if sum(v) / tot < 0.004:
        continue
    print(f"| {k} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / tot:.3f} | {sum(v) / len(v) / 1e3:.2f} |")
