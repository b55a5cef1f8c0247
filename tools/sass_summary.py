#!/usr/bin/env python
"""Per-kernel SASS summary of fluid_b200/libfluidb200.so -> profiles/<round>_sass_summary.md.

    python tools/sass_summary.py r02

Counts the mnemonics that show what the kernels are built from: bulk-copy TMA (UBLKCP), tensor-map TMA (UTMALDG),
mbarrier operations (SYNCS), named barriers (BAR), packed fp32 (FADD2 / FFMA2 / FMUL2, sm_100a), shared-memory and global
accesses, and the tensor-core families (none expected: nothing on this path is a contraction).  Static counts of the
whole kernel body (cold paths included), not executed instructions.  Stamped with the git head and the source stamp
bench.py uses for profiles/*_traffic.json."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "fluid_b200", "libfluidb200.so")
COLS = ["UBLKCP", "UTMALDG", "SYNCS", "BAR", "FADD2", "FFMA2", "FMUL2", "LDS.128", "LDS", "STS", "LDG", "STG", "SHFL", "HMMA/UTCMMA"]


def main(tag):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
    counts, total, name = collections.defaultdict(collections.Counter), collections.Counter(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if not m or name is None:
            continue
        op = m.group(1)
        total[name] += 1
        base = op.split(".")[0]
        c = counts[name]
        if base in ("UBLKCP", "UTMALDG", "SYNCS", "FADD2", "FFMA2", "FMUL2", "STS", "LDG", "STG", "SHFL"):
            c[base] += 1
        elif base == "BAR":
            c["BAR"] += 1
        elif base == "LDS":
            c["LDS.128" if ".128" in op else "LDS"] += 1
        elif base in ("HMMA", "UTCMMA", "UTCHMMA", "IMMA", "QGMMA", "HGMMA"):
            c["HMMA/UTCMMA"] += 1
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    import bench
    out = [f"# SASS summary of libfluidb200.so ({tag})", "",
           f"git head {head} (+ working tree), source stamp {bench.source_stamp()}; `cuobjdump -sass`, static counts per kernel "
           "(cold paths included).  UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, F*2 = packed fp32 (sm_100a).", "",
           "| kernel | instr | regs | " + " | ".join(COLS) + " |", "|---|---|---|" + "---|" * len(COLS)]
    import ctypes  # noqa: F401
    def demangle(n):
        try:
            return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
        except Exception:
            return n
    for n in sorted(total, key=lambda k: -total[k]):
        if total[n] < 40:
            continue
        out.append(f"| `{demangle(n)}` | {total[n]} | {regs.get(n, '')} | " + " | ".join(str(counts[n][c]) for c in COLS) + " |")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
