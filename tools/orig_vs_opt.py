#!/usr/bin/env python
"""BASELINE.json configs[2]: original-gcn vs optimize-gcn training epoch on the synthetic CiteSeer-shaped graph (10 % of the
edges between parties), with the REFERENCE'S OWN operator headers for both arms (oracle/_ref/gcn-original and gcn-optimize: the
reference's harness, engine and GCN operators compiled unchanged against cognn_b200/host/shim) on the CUDA library.

This is the API-faithful level-B path: every primitive call moves std::vector share matrices to the GPU and back, so the absolute
times are host-copy bound; what the comparison shows is the reference's own point -- the unoptimised operators push F = 3703-wide
rows through Scatter / Gather (and scale every edge row by two private normalisers), the optimised ones H = 16-wide rows.
Two parties by default: beyond two the reference's engine races on a loopback link (oracle/build_ref.py)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import refdrop  # noqa: E402
from tools import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="citeseer")
    ap.add_argument("--parties", type=int, default=2)
    ap.add_argument("--inter", type=float, default=0.1)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--mock", action="store_true", help="CPU mock of the C ABI (host-logic check only)")
    args = ap.parse_args()
    T = args.parties
    g = synth.make(args.shape, T, args.inter)
    rec = {"bench": "orig_vs_opt (configs[2])", "shape": args.shape, "parties": T, "inter_party_edges": g["inter_party_edges"], "N": g["N"],
           "E": g["E"], "cfg": {k: g["cfg"][k] for k in ("input_dim", "hidden_dim", "num_labels")}, "arms": {}}
    for name, binary, per_epoch in (("optimize-gcn", "gcn-optimize", 6), ("original-gcn", "gcn-original", 4)):
        out = None
        for attempt in range(4):
            rcs, o = refdrop.run(binary, g, T, per_epoch * args.epochs, args.mock, 35000 + 500 * (name == "original-gcn") + 40 * attempt,
                                 timeout=900)
            if all(len(x["loss"]) == args.epochs for x in o):
                out = o
                break
        if out is None:
            rec["arms"][name] = {"error": f"no complete run, return codes {rcs}", "tail": o[0]["tail"][-300:]}
            continue
        it = [max(o_["iteration_s"][k] for o_ in out) for k in range(per_epoch * args.epochs)]
        rec["arms"][name] = {"iterations_per_epoch": per_epoch, "iteration_s_max_over_parties": [round(x, 4) for x in it],
                             "epoch_s_last": round(sum(it[-per_epoch:]), 4), "loss_party0": out[0]["loss"],
                             "full_set_accuracy_party0": out[0]["acc_full"]}
    a = rec["arms"]
    if all("epoch_s_last" in a.get(k, {}) for k in ("optimize-gcn", "original-gcn")):
        rec["orig_over_opt_epoch_time"] = round(a["original-gcn"]["epoch_s_last"] / a["optimize-gcn"]["epoch_s_last"], 3)
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
