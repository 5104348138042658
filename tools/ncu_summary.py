"""Per-kernel averages of the metrics the roofline discussion uses, from an `ncu --set full` report (markdown table).
usage: ncu_summary.py report.ncu-rep"""
import collections, csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "duration us", 1e-3 if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else 1.0),
        ("dram__bytes_read.sum", "DRAM read MB", None), ("dram__bytes_write.sum", "DRAM write MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak", 1.0),
        ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %", 1.0),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory LSU wavefronts % of peak", 1.0),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory tensor-core wavefronts % of peak", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak", 1.0),
        ("launch__registers_per_thread", "registers / thread", 1.0),
        ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / CTA (KB)", None),
        ("smsp__inst_executed.sum", "warp instructions (M)", 1e-6)]
ki = hdr.index("Kernel Name")
groups = collections.OrderedDict()
for r in data:
    groups.setdefault(r[ki].split("(")[0].replace("void ", ""), []).append(r)
names = list(groups)
print("| metric | " + " | ".join(f"{n} ({len(groups[n])})" for n in names) + " |")
print("|---|" + "---|" * len(names))
for key, label, scale in want:
    if key not in hdr:
        continue
    i = hdr.index(key)
    u = units[i]
    cells = []
    for n in names:
        vals = [float(r[i].replace(",", "")) for r in groups[n] if r[i] not in ("", "n/a")]
        v = sum(vals) / max(len(vals), 1)
        if scale is None:
            b = u.split("/")[0]
            v = v / {"byte": 1e6, "Kbyte": 1e3, "Mbyte": 1.0, "Gbyte": 1e-3}.get(b, 1e6) if "MB" in label else v / {"byte": 1024.0, "Kbyte": 1.0}.get(b, 1024.0)
        else:
            v *= scale
        cells.append(f"{v:.1f}" if v < 1000 else f"{v:.0f}")
    print(f"| {label} | " + " | ".join(cells) + " |")
