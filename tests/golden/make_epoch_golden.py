"""Generates tests/golden/epoch_small.json: SHA-256 of every final weight share and of the whole message transcript of the
epoch ORACLE on two tiny fixed inputs (cora_small shape N=4 of build_from_source/config/cora_small_config.txt, and a
60-vertex 3-party graph).  A regression pin for the frozen semantics: any change to the protocol, the PRG stream layout,
truncation or encoding changes these digests.  Run here; the JSON is committed."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import epoch as ep  # noqa: E402
from tests.graphs import small_graph  # noqa: E402
from tools import synth  # noqa: E402


def digest(o):
    h = {}
    for p in range(o.T):
        for role, side in (("own", o.own), ("hlp", o.hlp)):
            for l in (0, 1):
                h[f"W{l}.{role}{p}"] = hashlib.sha256(np.ascontiguousarray(side[p]["W"][l]).tobytes()).hexdigest()
    t = hashlib.sha256()
    for it, src, dst, tag, data in o.msgs:
        t.update(f"{it}|{src}|{dst}|{tag}|".encode())
        t.update(np.ascontiguousarray(data).tobytes())
    h["transcript"] = t.hexdigest()
    h["n_messages"] = len(o.msgs)
    return h


def cases():
    g = synth.make("cora_small", 2)
    yield "cora_small_2p", g["edges"], g["tid"], 2, g["feats"], g["labels"], g["cfg"]
    g = small_graph(n=60, n_edges=220, F=12, C=4, T=3, seed=3)
    cfg = dict(input_dim=12, hidden_dim=8, num_labels=4, learning_rate=0.5, train_ratio=0.4, val_ratio=0.2)
    yield "random60_3p", g["edges"], g["tid"], 3, g["feats"], g["labels"], cfg


if __name__ == "__main__":
    out = {}
    for name, edges, tid, T, feats, labels, cfg in cases():
        o = ep.EpochOracle(edges, tid, T, feats, labels, cfg)
        o.run(6)
        out[name] = digest(o)
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "epoch_small.json"), "w"), indent=1)
    print({k: v["transcript"][:16] for k, v in out.items()})
