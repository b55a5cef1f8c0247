"""Split one launch of an ncu report into regions of its SASS (between the SYNCS hand-off instructions)
and print executed instructions, stall samples and the main stall reasons per region.
usage: python tools/ncu_sass_regions.py report.ncu-rep [launch_index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; k = sys.argv[2] if len(sys.argv) > 2 else "0"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", k,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
body = []
for r in rows[hi + 1:]:          # newer ncu prints the table twice (per view): keep the first
    if r and r[0] in ("Address", "Kernel Name"):
        break
    if len(r) > col["Instructions Executed"]:
        body.append(r)
tot_i = sum(int(r[col["Instructions Executed"]]) for r in body)
tot_s = sum(int(r[col["# Samples"]]) for r in body)
print("instructions", tot_i, "samples", tot_s)
reg = []; cur = {"start": 0, "n": 0, "inst": 0, "smp": 0, "st": {}, "sync": ""}
for idx, r in enumerate(body):
    src = r[col["Source"]].strip()
    cur["n"] += 1
    cur["inst"] += int(r[col["Instructions Executed"]]); cur["smp"] += int(r[col["# Samples"]])
    for s in stalls:
        v = int(r[col[s]] or 0)
        if v: cur["st"][s] = cur["st"].get(s, 0) + v
    if "SYNCS.ARRIVE" in src or "EXIT" in src.split()[0:2] or src.startswith("EXIT") or "SYNCS.PHASECHK" in src:
        cur["sync"] = src[:60]; cur["end"] = idx
        reg.append(cur); cur = {"start": idx + 1, "n": 0, "inst": 0, "smp": 0, "st": {}, "sync": ""}
reg.append(cur)
for c in reg:
    if c["inst"] < tot_i * 0.002 and c["smp"] < tot_s * 0.002: continue
    top = sorted(c["st"].items(), key=lambda x: -x[1])[:4]
    print(f"sass {c['start']:5d}+{c['n']:4d} inst {100*c['inst']/tot_i:5.1f}% smp {100*c['smp']/tot_s:5.1f}%  ends: {c['sync']:<58} "
          + " ".join(f"{k[6:]}={v}" for k, v in top))
