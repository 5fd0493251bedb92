"""Generates tests/golden/ideal_softmax.json: inputs and outputs of the prediction-layer stand-in (orc_ideal_softmax /
cgb_ideal_softmax) on a small fixed case, plus orc_det_exp at a handful of points as IEEE-754 bit patterns.  A regression
pin of the restated exp: the C oracle, its numpy restatement and the CUDA kernel must all reproduce these bits.  Run here;
the JSON is committed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import pyoracle as po  # noqa: E402


def case():
    rng = np.random.default_rng(2024)
    n, C, f = 24, 7, 16
    z = (rng.normal(0, 3, size=(n, C)) * (1 << f)).astype(np.int64)
    z[0] = 12345          # constant row
    z[1, 3] = 700 << f    # dominant class, the rest underflow
    z1 = rng.integers(0, 1 << 64, size=(n, C), dtype=np.uint64)
    z0 = z.view(np.uint64) - z1
    labels = rng.integers(0, C, size=n).astype(np.int32)
    return z0, z1, labels, 10, f


EXP_POINTS = [0.0, -1e-300, -0.5, -1.0, -0.6931471805599453, -10.25, -37.0, -100.0, -699.9, -700.0, -700.1, -745.2]

if __name__ == "__main__":
    po.build()
    z0, z1, labels, train, f = case()
    P, pmy = po.ideal_softmax(z0, z1, labels, train, f)
    out = {"z0": z0.ravel().tolist(), "z1": z1.ravel().tolist(), "labels": labels.tolist(), "train_rows": train, "f": f,
           "shape": list(z0.shape), "P": P.ravel().tolist(), "pmy": pmy.ravel().tolist(),
           "det_exp": {repr(x): int(np.float64(po.det_exp(x)).view(np.uint64)) for x in EXP_POINTS}}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ideal_softmax.json"), "w"))
    print("rows", z0.shape, "sum P row0", int(P[0].sum()))
