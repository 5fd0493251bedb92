"""Pins oracle/graph_index.py against the worked example of SURVEY.md 3.6 (hand-derived from
include/ss_vertex_centric_algo_kernel.h:295-534 of the reference)."""
import numpy as np

from oracle import graph_index as gi

# cora_small shape: N=4, undirected pairs 0-1, 1-2, 2-3, 0-2 stored in both directions, partition vid % 2
EDGES = [(0, 1), (1, 0), (1, 2), (2, 1), (2, 3), (3, 2), (0, 2), (2, 0)]
TID = [0, 1, 0, 1]


def L(a):
    return [int(v) for v in a]


def test_worked_example_party0():
    tiles, ivs = gi.build_all(EDGES, TID, 2, no_dummy_edge=True)
    t0, iv = tiles[0], ivs[0]
    assert [tuple(e) for e in t0.edges.tolist()] == [(0, 1), (0, 2), (2, 0), (2, 1), (2, 3)]
    assert L(iv["localVertexPos"]) == [0, 2]
    assert L(iv["localVertexInDeg"]) == [2, 3]
    assert L(iv["updateSrcVertexPos"][0]) == [2, 0] and L(iv["updateDstVertexPos"][0]) == [0, 2]
    assert L(iv["updateSrcOutDeg"][0]) == [3, 2] and L(iv["updateDstInDeg"][0]) == [2, 3]
    assert L(iv["isGatherDstVertexDummy"][0]) == [0, 0]
    assert L(iv["updateSrcVertexPos"][1]) == [0, 2, 2] and L(iv["updateDstVertexPos"][1]) == [1, 1, 3]
    assert L(iv["updateSrcOutDeg"][1]) == [2, 3, 3] and L(iv["updateDstInDeg"][1]) == [0, 0, 0]
    assert L(iv["remoteMirrorVertexPos"][1]) == [0, 2, 2]
    assert L(iv["remoteUpdateDstInDeg"][1]) == [2, 3, 3]
    assert L(iv["isGatherDstVertexDummy"][1]) == [0, 0]
    # what party 1 receives from party 0
    assert L(ivs[1]["remoteMirrorVertexPos"][0]) == [1, 1, 3]


def test_dummy_self_edge_increments_both_degrees():
    # vertex 2 (party 0) has only a remote in-edge: gets a dummy self edge, inDeg and outDeg + 1 (ssk.h:412-418)
    edges = [(0, 1), (1, 0), (1, 2), (2, 1)]
    tid = [0, 1, 0, 1]
    tiles, ivs = gi.build_all(edges, tid, 2, no_dummy_edge=True)
    iv = ivs[0]
    assert L(iv["localVertexPos"]) == [0, 2]
    # both local vertices have no LOCAL in-edge -> both get dummies
    assert L(iv["updateSrcVertexPos"][0]) == [0, 2] and L(iv["isUpdateSrcVertexDummy"][0]) == [1, 1]
    assert L(iv["isGatherDstVertexDummy"][0]) == [1, 1]
    assert L(iv["localVertexInDeg"]) == [2, 2]  # 1 remote in-edge + 1 dummy
    assert L(iv["updateSrcOutDeg"][1]) == [2, 2]  # 1 real out-edge + 1 dummy


def test_power_of_two_padding_without_r_flag():
    tiles, ivs = gi.build_all(EDGES, TID, 2, no_dummy_edge=False)
    iv = ivs[0]
    # vertex 2 of party 0 has 1 local source (0) -> stays 1; vertex 0 has 1 (2) -> stays 1
    assert L(iv["updateSrcVertexPos"][0]) == [2, 0]
    # mirror 1 has sources [0, 2] (2 = pow2), mirror 3 has [2]
    assert L(iv["updateSrcVertexPos"][1]) == [0, 2, 2]
    edges = EDGES + [(0, 3), (3, 0), (2, 5), (5, 2), (4, 5), (5, 4)]
    tid = [0, 1, 0, 1, 0, 1]
    tiles, ivs = gi.build_all(edges, tid, 2, no_dummy_edge=False)
    iv = ivs[0]
    # mirror 5 has sources [2, 4]; mirror 3 has [0, 2]; mirror 1 has [0, 2]: all pow2.  local vertex 4: no local source
    # -> padded to one dummy self entry without degree change (ssk.h:369-374)
    assert L(iv["localVertexPos"]) == [0, 2, 4]
    assert L(iv["updateSrcVertexPos"][0]) == [2, 0, 4]
    assert L(iv["isUpdateSrcVertexDummy"][0]) == [0, 0, 1]
    assert L(iv["localVertexInDeg"]) == [3, 4, 1]
    for t in range(2):
        dst = iv["updateDstVertexPos"][t]
        _, counts = np.unique(dst, return_counts=True)
        assert all(c & (c - 1) == 0 for c in counts)


def test_csr_from_pos_matches_adjacency():
    tiles, ivs = gi.build_all(EDGES, TID, 2)
    iv = ivs[0]
    rowptr, col = gi.csr_from_pos(iv["updateSrcVertexPos"][0], iv["updateDstVertexPos"][0], iv["localVertexPos"],
                                  iv["localVertexPos"])
    assert L(rowptr) == [0, 1, 2] and L(col) == [1, 0]
    rowptr, col = gi.csr_from_pos(iv["updateSrcVertexPos"][1], iv["updateDstVertexPos"][1], iv["localVertexPos"],
                                  ivs[1]["localVertexPos"])
    assert L(rowptr) == [0, 2, 3] and L(col) == [0, 1, 1]
