"""Compare the raw float32 dumps of go/cmd/dump (the real reference) with the committed goldens, bit for bit.

    python tests/golden/compare_go_dump.py <dump_root> <golden_dir>

<dump_root>/<case>/<field>_<step>.f32 against <golden_dir>/<case>.npz[<field>_<step>].  +0 and -0 compare equal
(the comparison the parity tests use); anything else must be identical.  Prints max-abs and relative-L2 per field
where they differ and exits 1."""
import os
import sys

import numpy as np


def main(dump_root, golden_dir):
    bad = 0
    for case in sorted(os.listdir(dump_root)):
        gold_path = os.path.join(golden_dir, case + ".npz")
        if not os.path.exists(gold_path):
            print(f"{case}: no golden")
            continue
        gold = np.load(gold_path)
        for key in sorted(gold.files):
            path = os.path.join(dump_root, case, key + ".f32")
            if not os.path.exists(path):
                print(f"{case}:{key}: not dumped")
                bad += 1
                continue
            ref = np.fromfile(path, dtype="<f4").reshape(gold[key].shape)
            if np.array_equal(ref, gold[key]):
                continue
            d = np.abs(ref.astype(np.float64) - gold[key].astype(np.float64))
            rel = np.sqrt((d ** 2).sum() / max((ref.astype(np.float64) ** 2).sum(), 1e-300))
            print(f"{case}:{key}: DIFFERS max_abs={d.max():.3e} rel_l2={rel:.3e} cells={int((d > 0).sum())}")
            bad += 1
        print(f"{case}: checked {len(gold.files)} arrays")
    print("go reference == goldens, bit for bit" if not bad else f"{bad} arrays differ from the Go reference")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
