"""Markdown table of the kernels in an ncu report (raw page): python tools/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[0]
cols = [("Kernel Name", "kernel", str), ("gpu__time_duration.sum", "time_us", float), ("dram__bytes_read.sum", "dram_read_MB", float),
        ("dram__bytes_write.sum", "dram_write_MB", float), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", float),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct", float),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_pipe_pct", float),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct", float),
        ("smsp__inst_executed.sum", "warp_inst", float), ("launch__registers_per_thread", "regs", float),
        ("launch__grid_size", "grid", float), ("launch__block_size", "block", float),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem_KB", float), ("lts__t_sector_hit_rate.pct", "l2_hit_pct", float)]
ix = [(h.index(c), n, t) for c, n, t in cols if c in h]
units = rows[1]
print("| " + " | ".join(n for _, n, _ in ix) + " |")
print("|" + "---|" * len(ix))
seen = set()
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "")
    if name in seen:
        continue
    seen.add(name)
    out = []
    for i, n, t in ix:
        v = r[i]
        if t is str:
            out.append(name)
            continue
        try:
            f = float(v)
        except ValueError:
            out.append(v); continue
        u = units[i]
        if n == "time_us" and u in ("ms", "msecond"): f *= 1e3
        if n == "time_us" and u in ("ns", "nsecond"): f /= 1e3
        if n.endswith("_MB") and u == "Gbyte": f *= 1e3
        if n.endswith("_MB") and u == "Kbyte": f /= 1e3
        if n == "dyn_smem_KB" and u == "byte": f /= 1024
        out.append(f"{f:,.0f}" if n in ("warp_inst", "grid", "block", "regs") else f"{f:,.1f}")
    print("| " + " | ".join(out) + " |")
