"""Opcode histogram per kernel of the built library (cuobjdump -sass): which kernels carry tcgen05 / TMA / bulk-copy /
asynchronous-copy instructions.  usage: sass_histogram.py > profiles/<round>_sass_histogram.md"""
import collections, os, re, subprocess, sys
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(here, "scann_b200", "libscann_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "LDGSTS", "SYNCS", "LDG", "STG", "RED",
        "ATOMS", "LDS", "STS", "MUFU", "FFMA", "SHFL", "BAR"]
hist = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        hist.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        op = m.group(1)
        for c in cols:
            if op == c or op.startswith(c + ".") or (c == "UTCHMMA" and op.startswith("UTCHMMA")):
                hist[name][c] += 1
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
for k in sorted(hist):
    if k.startswith("void at::") or "pipe_give_up" in k:
        continue
    print(f"| {k[:46]} | " + " | ".join(str(hist[k][c]) for c in cols) + " |")
