"""The C++ index-vector builder (cognn_b200/host, ssk.h:295-534 with -r 1) against the numpy oracle and the SURVEY 3.6
worked example.  Host only: runs without a GPU."""
import numpy as np
import pytest

from cognn_b200 import engine as eng
from oracle import epoch as ep
from oracle import graph_index as gi
from tests.graphs import small_graph


def test_worked_example_party0():
    edges = [(0, 1), (1, 0), (1, 2), (2, 1), (2, 3), (3, 2), (0, 2), (2, 0)]
    g = eng.build_party_graph(edges, [0, 1, 0, 1], 2, 0)
    assert g["vids"].tolist() == [0, 2] and g["in_deg"].tolist() == [2, 3] and g["offsets"].tolist() == [0, 2, 4]
    # rows: party 0's vertices (0, 2), then party 1's (1, 3); columns are local rows of party 0 (0 -> v0, 1 -> v2)
    assert g["rowptr"].tolist() == [0, 1, 2, 4, 5]
    assert g["col"].tolist() == [1, 0, 0, 1, 1]


@pytest.mark.parametrize("T,partition", [(2, "mod"), (3, "mod"), (4, "block"), (5, "mod")])
def test_matches_numpy_oracle(T, partition):
    g = small_graph(n=90, n_edges=400, F=3, C=3, T=T, seed=10 + T, partition=partition)
    tiles, ivs = gi.build_all(g["edges"], g["tid"], T, no_dummy_edge=True)
    for me in range(T):
        got = eng.build_party_graph(g["edges"], g["tid"], T, me)
        iv = ivs[me]
        assert np.array_equal(got["vids"], iv["localVertexPos"])
        assert np.array_equal(got["in_deg"], iv["localVertexInDeg"])
        assert np.array_equal(got["in_deg_raw"], tiles[me].in_deg[iv["localVertexPos"].astype(np.int64)].astype(np.uint64))
        rowptr, col = ep.fused_csr(iv, ivs)
        assert np.array_equal(got["rowptr"], rowptr) and np.array_equal(got["col"], col)


def test_engine_header_symbols_exported():
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "cognn_b200_engine.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(cge_[a-z0-9_]+)\s*\(", src)))
    h = eng.load_host()
    assert sorted(eng.ENGINE_SYMBOLS) == names
    for n in names:
        assert hasattr(h, n)
