"""Pins the epoch oracle (oracle/epoch.py): reconstructed values against a float64 GCN of the same dataflow, glibc
weights against a committed golden vector, share/message invariants."""
import json
import os

import numpy as np
import pytest

from oracle import epoch as ep
from tests.graphs import small_graph

HERE = os.path.dirname(os.path.abspath(__file__))


def test_glibc_weight_init_golden():
    gold = json.load(open(os.path.join(HERE, "golden", "glibc_rand42.json")))
    W = ep.glibc_init_weight(2, 3)
    assert np.allclose(W.ravel(), gold["initWeight_2x3"], rtol=0, atol=0)
    # srand(42) is re-seeded per layer (gcn.h:841): both layers start from the same rand() sequence
    W2 = ep.glibc_init_weight(3, 3)
    lim1, lim2 = np.sqrt(6.0 / 5), np.sqrt(6.0 / 6)
    assert np.allclose((W.ravel()[:3] + lim1) / lim1, (W2.ravel()[:3] + lim2) / lim2)


@pytest.mark.parametrize("T", [2, 3, 4])
def test_epoch_matches_float64_gcn(T):
    g = small_graph(n=60, n_edges=220, F=12, C=4, T=T, seed=T)
    cfg = dict(input_dim=12, hidden_dim=8, num_labels=4, learning_rate=0.5, train_ratio=0.4, val_ratio=0.2)
    o = ep.EpochOracle(g["edges"], g["tid"], T, g["feats"], g["labels"], cfg)
    o.run(6)
    ref = ep.PlainGCN(o, g["feats"], g["labels"])
    ref.epoch(6)
    for p in range(T):
        for layer in (0, 1):
            got = o.weights_plain(p, layer)
            assert np.allclose(got, ref.W[p][layer], atol=2e-2), (p, layer, np.abs(got - ref.W[p][layer]).max())
    # FedAvg leaves every replica equal (gcn.h:747-802)
    for p in range(1, T):
        for layer in (0, 1):
            assert np.array_equal(o.own[p]["W"][layer] + o.hlp[p]["W"][layer], o.own[0]["W"][layer] + o.hlp[0]["W"][layer])
    # inference = iterations 0 and 1: opened predictions match the float softmax
    o2 = ep.EpochOracle(g["edges"], g["tid"], T, g["feats"], g["labels"], cfg)
    o2.run(2)
    ref2 = ep.PlainGCN(o2, g["feats"], g["labels"])
    ref2.epoch(2)
    assert len(o2.log) == T
    for p in range(T):
        lg = [m for m in o2.log if m["party"] == p][0]
        pred_ref = ref2.probs[p].argmax(axis=1)
        acc_ref = float((pred_ref == ref2.labels[p]).mean())
        assert abs(lg["acc_full"] - acc_ref) <= 2.0 / max(1, o2.n[p])


def test_message_schedule_two_parties_has_no_update_blocks():
    g = small_graph(n=30, n_edges=90, F=6, C=3, T=2, seed=5)
    cfg = dict(input_dim=6, hidden_dim=4, num_labels=3, learning_rate=0.5, train_ratio=0.4, val_ratio=0.2)
    o = ep.EpochOracle(g["edges"], g["tid"], 2, g["feats"], g["labels"], cfg)
    o.run(6)
    tags = {m[3].split(".")[0] for m in o.msgs}
    assert "upd" not in tags and "w0" not in tags  # T = 2: the primary helper of the other party is the sender itself
    assert {"mm0", "mm1", "rm0", "rm1", "om", "res0", "open_p"} <= tags
    keys = [m[:4] for m in o.msgs]
    assert len(keys) == len(set(keys))  # (iteration, src, dst, tag) identifies a message
    # determinism: same seeds, same bytes
    o2 = ep.EpochOracle(g["edges"], g["tid"], 2, g["feats"], g["labels"], cfg)
    o2.run(6)
    assert len(o.msgs) == len(o2.msgs)
    for a, b in zip(o.msgs, o2.msgs):
        assert a[:4] == b[:4] and np.array_equal(a[4], b[4])


def test_epoch_digests_match_committed_golden():
    """Regression pin of the frozen semantics (tests/golden/epoch_small.json, made by make_epoch_golden.py)."""
    from tests.golden import make_epoch_golden as mk

    gold = json.load(open(os.path.join(HERE, "golden", "epoch_small.json")))
    for name, edges, tid, T, feats, labels, cfg in mk.cases():
        o = ep.EpochOracle(edges, tid, T, feats, labels, cfg)
        o.run(6)
        assert mk.digest(o) == gold[name], name
