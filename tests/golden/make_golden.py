"""Regenerates tests/golden/*.npz from the fp64 NumPy oracle (python tests/golden/make_golden.py).
The reference ships no golden vectors and cannot run here, so these files
freeze the ORACLE's answers: later refactors of the oracle or the kernels are pinned to them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cases  # noqa: E402

SEED = 2018


def main():
    for name in ["A", "C", "U", "M", "MU", "B"]:
        case = cases.pair_case(name, seed=SEED)
        o = cases.oracle_eval(case)
        out = {"logits": o["logits"].astype(np.float64), "loss": np.asarray(o["loss"], np.float64)}
        for k, v in o["grads"].items():
            out["grad:" + k] = v.astype(np.float32)
        np.savez_compressed(os.path.join(HERE, "pair_%s.npz" % name), **out)
        print(name, out["logits"].shape, float(out["loss"]))


if __name__ == "__main__":
    main()
