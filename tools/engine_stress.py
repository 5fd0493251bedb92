#!/usr/bin/env python
"""Repeat the small engine-vs-oracle comparison many times in one process and report every mismatch (which share, which
tensor, how many words): a race between the engine's streams shows up as an occasional mismatch that a single test run
misses.  python tools/engine_stress.py [--reps 60] [--record 1]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import epoch as ep  # noqa: E402  (test infrastructure: the checker)
from tests.graphs import small_graph  # noqa: E402
from tests.test_gpu_engine import NAMES, oracle_tensor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=60)
    ap.add_argument("--record", type=int, default=1)
    ap.add_argument("--iters", type=int, default=6)
    args = ap.parse_args()
    from cognn_b200 import engine as eng

    bad = 0
    for T in (2, 3, 4):
        g = small_graph(n=70, n_edges=260, F=10, C=4, T=T, seed=40 + T)
        cfg = dict(input_dim=10, hidden_dim=8, num_labels=4, learning_rate=0.5, train_ratio=0.4, val_ratio=0.2)
        o = ep.EpochOracle(g["edges"], g["tid"], T, g["feats"], g["labels"], cfg)
        o.run(args.iters)
        want_msgs = {(m[0], m[1], m[2], m[3]): m[4] for m in o.msgs}
        for rep in range(args.reps):
            e = eng.Engine(T, cfg, record=bool(args.record))
            e.load(g["edges"], g["tid"], g["feats"], g["labels"])
            e.run(args.iters)
            for owner in range(T):
                for role in (0, 1):
                    for name in NAMES:
                        want = oracle_tensor(o, owner, role, name)
                        got = e.download(owner, role, name)
                        if got.shape != want.shape or not np.array_equal(got, want):
                            bad += 1
                            nd = int((got != want).sum()) if got.shape == want.shape else -1
                            print(f"T={T} rep={rep} MISMATCH owner={owner} role={role} {name} words={nd}/{want.size}", flush=True)
            if args.record:
                first = None
                for m in e.messages():
                    if m[3].startswith("setup"):
                        continue
                    k = (m[0], m[1], m[2], m[3])
                    if k in want_msgs and not np.array_equal(m[4], want_msgs[k]):
                        first = first or k
                if first:
                    print(f"T={T} rep={rep} first differing message: {first}", flush=True)
            e.close()
    print(f"done: {bad} mismatching tensors")


if __name__ == "__main__":
    main()
