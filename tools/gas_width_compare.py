#!/usr/bin/env python
"""BASELINE configs[2] data-path comparison: original-gcn moves F-wide rows through Scatter/Gather (GAS widths per
epoch [F, H, -, H], original-gcn/gcn.h:807-851) where CoGNN-Opt moves [H, C, C, H] (optimize-gcn/gcn.h:898-948).
Times the fused gather-sum of every party of the 4-party CiteSeer-shaped graph (10 % inter-party edges) at both width
sets.  Only the share-gather data path is compared: original-gcn's fused NN primitives (twoPartyGCNForwardNN, ...BackwardNN)
are absent from the reference tree and their operator arm is not rebuilt this round."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import cognn_b200  # noqa: E402
from cognn_b200 import engine as eng  # noqa: E402
from tools import synth  # noqa: E402


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "citeseer"
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    g = synth.make(shape, T, 0.1 if shape == "citeseer" else None)
    F, H, C = g["cfg"]["input_dim"], g["cfg"]["hidden_dim"], g["cfg"]["num_labels"]
    ctx = cognn_b200.Context(0)
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    res = {"orig": {}, "opt": {}}
    widths = {"orig": [F, H, H], "opt": [H, C, C, H]}
    for p in range(T):
        pg = eng.build_party_graph(g["edges"], g["tid"], T, p)
        n = pg["vids"].size
        csr = ctx.csr_create(torch.from_numpy(pg["rowptr"].view(np.int32)).to(dev), torch.from_numpy(pg["col"].view(np.int32)).to(dev), n)
        for arm, ws in widths.items():
            for D in sorted(set(ws)):
                x = torch.randint(-2**63, 2**63 - 1, (n, D), dtype=torch.int64, device=dev, generator=gen)
                y = torch.empty((csr.n_rows, D), dtype=torch.int64, device=dev)
                for _ in range(3):
                    ctx.gather_sum(csr, x, None, out=y)
                ts = []
                for _ in range(7):
                    flush.fill_(1)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); ctx.gather_sum(csr, x, None, out=y); b.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
                ts.sort()
                res[arm].setdefault(D, []).append({"party": p, "edges": int(pg["col"].size), "ms": ts[len(ts) // 2],
                                                   "alg_bytes": bench.algorithmic_bytes(csr.n_rows, pg["col"].size, D)})
        csr.destroy()
    out = {"bench": "gas_width_compare", "shape": shape, "parties": T, "F": F, "H": H, "C": C, "widths": widths}
    for arm, ws in widths.items():
        ms = sum(max(r["ms"] for r in res[arm][D]) for D in ws)       # parties run in parallel: max over parties
        byts = sum(sum(r["alg_bytes"] for r in res[arm][D]) for D in ws)
        out[arm] = {"gas_ms_per_epoch": ms, "alg_bytes_per_epoch_all_parties": byts}
    out["bytes_ratio_orig_over_opt"] = out["orig"]["alg_bytes_per_epoch_all_parties"] / out["opt"]["alg_bytes_per_epoch_all_parties"]
    out["time_ratio_orig_over_opt"] = out["orig"]["gas_ms_per_epoch"] / out["opt"]["gas_ms_per_epoch"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
