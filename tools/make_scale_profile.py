"""gpurun_out/r02_scale/ (tools/r02_multigpu.sh N on N-GPU boxes) -> profiles/r02_scale.md, profiles/r02_bench_karman4096_<N>gpu.json,
profiles/r02_multi_gpu_check_<N>gpu.log.   usage: python tools/make_scale_profile.py [round-tag]"""
import json, os, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
SRC = os.path.join(ROOT, "gpurun_out", f"{TAG}_scale")
DST = os.path.join(ROOT, "profiles")


def last_json(path):
    with open(path) as fh:
        lines = [ln for ln in fh.read().splitlines() if ln.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def main():
    rows, base = [], None
    for n in (1, 2, 4, 8):
        p = os.path.join(SRC, f"bench_karman4096_{n}gpu.json")
        if not os.path.exists(p):
            continue
        d = last_json(p)
        if d is None:
            continue
        with open(os.path.join(DST, f"{TAG}_bench_karman4096_{n}gpu.json"), "w") as fh:
            fh.write(json.dumps(d) + "\n")
        q = os.path.join(SRC, f"bench_karman4096_{n}gpu_nooverlap.json")
        nov = last_json(q) if os.path.exists(q) else None
        log = os.path.join(SRC, f"multi_gpu_check_{n}gpu.log")
        if os.path.exists(log):
            shutil.copy(log, os.path.join(DST, f"{TAG}_multi_gpu_check_{n}gpu.log"))
        if n == 1:
            base = d
        rows.append((n, d, nov))
    out = [f"# Weak scaling over row slabs, round {TAG} (Karman 4098^2 cells per GPU, BFECC + confinement, developed flow; `tools/r02_multigpu.sh N`)", "",
           "`bench.py --gpus N --steps 20 --warmup 5` under torchrun, exactly the driver's invocation; the line carries the N-rank vs 1-GPU "
           "parity check (`parity_check`) and the config[3] leg (`config3_jet16384`).  Efficiency = (value / N) / value(N = 1).", "",
           "| N | ms per step | G cell-steps/s | efficiency | exchange in front of the step (ms) | e2e display loop G (eff.) | full-field loop G | "
           "jet 16384^2 per GPU: ms, G (eff.) | parity_check | clocks (MHz, samples, reasons) |", "|---|---|---|---|---|---|---|---|---|---|"]
    for n, d, nov in rows:
        eff = d["value"] / n / base["value"] if base else float("nan")
        e2e = d.get("e2e") or {}
        e2e_eff = (e2e.get("value", 0) / n / base["e2e"]["value"]) if base and base.get("e2e") and e2e else float("nan")
        c3, b3 = d.get("config3_jet16384"), base.get("config3_jet16384") if base else None
        c3s = f"{c3['ms_per_step']:.3f}, {c3['value'] / 1e9:.1f} ({c3['value'] / n / b3['value']:.3f})" if c3 and b3 else "-"
        pc = d.get("parity_check")
        pcs = "-" if n == 1 else (f"ok={pc['ok']} max|diff|={max(pc['max_abs_diff'].values())}" if pc else "missing")
        ck = d.get("clocks") or {}
        out.append(f"| {n} | {d['ms_per_step']:.4f} | {d['value'] / 1e9:.2f} | {eff:.3f} | {nov['ms_per_step']:.4f} |" if nov else
                   f"| {n} | {d['ms_per_step']:.4f} | {d['value'] / 1e9:.2f} | {eff:.3f} | - |")
        out[-1] += (f" {e2e.get('value', 0) / 1e9:.2f} ({e2e_eff:.3f}) | {(e2e.get('full_field_loop') or {}).get('value', 0) / 1e9:.2f} | {c3s} | {pcs} | "
                    f"{ck.get('sm_mhz')}, {ck.get('samples')}, {ck.get('reasons')} |")
    out += ["", "`r02_multi_gpu_check_<N>gpu.log`: `tests/multi_gpu_check.py` on N ranks (three presets, explicit and overlapped exchange, peer memory "
            "and NCCL transports, projection only): every field of every case bit-identical to the single-GPU run."]
    with open(os.path.join(DST, f"{TAG}_scale.md"), "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
