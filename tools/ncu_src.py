"""Per-source-line instruction counts of one launch in an ncu report.
usage: python tools/ncu_src.py report.ncu-rep launch_index [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; k = sys.argv[2]; N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", k,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if len(r) > 8 and r[0] == 'Line No')
hdr = rows[hi]
iE = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
out = []; tot_i = tot_s = 0
print(rows[1][1][:80])
for r in rows[hi + 1:]:
    if len(r) <= iE or not r[0].isdigit():
        continue
    try:
        n = int(r[iE]); s = int(r[iS])
    except Exception:
        continue
    tot_i += n; tot_s += s
    out.append((n, s, r[0], r[1]))
print("total warp instr", tot_i, "samples", tot_s)
for n, s, ln, src in sorted(out, key=lambda x: -x[0])[:N]:
    print(f"{n:>10} {100*n/max(tot_i,1):5.1f}% smp {100*s/max(tot_s,1):5.1f}% | {ln:>4} {src.strip()[:120]}")
