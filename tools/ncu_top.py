"""Top stalled SASS instructions of an ncu report: python tools/ncu_top.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
out = []; tot = 0
for k, r in enumerate(data):
    try: n = int(r[ix['# Samples']])
    except Exception: continue
    tot += n; out.append((n, k, r))
print("total samples", tot)
if len(sys.argv) > 3:      # address range dump
    a0, a1 = sys.argv[3], sys.argv[4]
    for n, k, r in out:
        a = r[ix['Address']][-5:]
        if a0 <= a <= a1:
            print(a, str(n).rjust(5), r[ix['Instructions Executed']].rjust(8), r[ix['Source']][:90])
    sys.exit()
for n, k, r in sorted(out, key=lambda x: -x[0])[:N]:
    st = {c: int(r[ix[c]]) for c in stall_cols if r[ix[c]] not in ('', '0')}
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(n, r[ix['Address']][-5:], r[ix['Instructions Executed']].rjust(8), r[ix['Source']][:70], top)
