"""Device graph ingest / index-vector builder (cgb_party_graph_build) against the numpy oracle
(oracle/graph_index.py, the restatement of graph_io_util.h:40-208 + ssk.h:295-534) and the C++ host builder."""
import numpy as np
import pytest

from oracle import graph_index as gi
from tests.graphs import small_graph
from tests.util import rand_u64, to_dev, to_np

pytestmark = pytest.mark.gpu


def check_against_host_builder(cgb, edges, tid, T):
    from cognn_b200 import engine as eng

    edges = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1, 2)
    tid = np.ascontiguousarray(tid, dtype=np.int64)
    for me in range(T):
        want = eng.build_party_graph(edges, tid, T, me)
        for src in ("host", "device"):
            import torch

            e = torch.from_numpy(edges)
            t = torch.from_numpy(tid)
            if src == "device":
                e, t = e.cuda(), t.cuda()
            got = cgb.party_graph_build(e, t, T, me)
            assert got["offsets"] == want["offsets"].tolist(), (me, src)
            assert np.array_equal(to_np(got["vids"]), want["vids"]), (me, src)
            assert np.array_equal(to_np(got["in_deg_raw"]), want["in_deg_raw"]), (me, src)
            assert np.array_equal(to_np(got["in_deg"]), want["in_deg"]), (me, src)
            assert np.array_equal(to_np(got["rowptr"]), want["rowptr"]), (me, src)
            assert np.array_equal(to_np(got["col"]), want["col"]), (me, src)
            # isLocalVertexBorder (graph_io_util.h:169): a local vertex with at least one out-edge into another party
            border = np.zeros(int(got["n_local"]), dtype=np.uint8)
            if edges.size:
                out_remote = (tid[edges[:, 0]] == me) & (tid[edges[:, 1]] != me)
                local_of = np.cumsum(tid == me) - 1
                border[local_of[edges[out_remote, 0]]] = 1
            assert np.array_equal(to_np(got["is_border"]), border), (me, src)
            yield me, got, want


@pytest.mark.parametrize("T,partition", [(1, "mod"), (2, "mod"), (3, "block"), (4, "mod"), (8, "mod"), (16, "mod")])
def test_ingest_matches_host_builder_and_oracle(cgb, oracle, T, partition):
    g = small_graph(n=300, n_edges=2000, F=4, C=3, T=T, seed=10 + T, partition=partition, multi_edges=40, isolated=7)
    edges, tid = np.asarray(g["edges"], dtype=np.int64), np.asarray(g["tid"], dtype=np.int64)
    tiles, ivs = gi.build_all(edges, tid, T)
    rng = np.random.default_rng(T)
    for me, got, want in check_against_host_builder(cgb, edges, tid, T):
        iv = ivs[me]
        assert np.array_equal(to_np(got["vids"]).astype(np.int64), iv["localVertexPos"].astype(np.int64))
        assert np.array_equal(to_np(got["in_deg"]).astype(np.int64), iv["localVertexInDeg"].astype(np.int64))
        # the CSR that came out of the device ingest drives the gather kernel: same sums as the oracle on the host arrays
        x = rand_u64(rng, got["n_local"], 5)
        y = to_np(cgb.gather_sum(got["csr"], to_dev(x)))
        assert np.array_equal(y, oracle.gather_sum_csr(want["rowptr"], want["col"], x))
        got["csr"].destroy()


def test_ingest_edge_cases(cgb):
    import torch

    # no edges at all; a party without vertices; self loops and repeated edges; every edge crossing parties
    cases = [
        (np.zeros((0, 2), dtype=np.int64), np.array([0, 1, 0, 1]), 2),
        (np.array([[0, 1], [1, 0], [2, 2], [2, 2], [1, 2]]), np.array([0, 0, 0]), 2),
        (np.array([[0, 1], [1, 0], [2, 3], [3, 2], [0, 3], [3, 0]]), np.array([0, 1, 0, 1]), 2),
    ]
    for edges, tid, T in cases:
        list(check_against_host_builder(cgb, edges, tid, T))
    with pytest.raises(Exception):
        cgb.party_graph_build(torch.tensor([[0, 9]]), torch.tensor([0, 1]), 2, 0)  # vertex id out of range
    with pytest.raises(Exception):
        cgb.party_graph_build(torch.tensor([[0, 1]]), torch.tensor([0, 2]), 2, 0)  # tile id out of range


def test_ingest_at_scale_properties(cgb):
    """3M-edge power-law graph, 4 parties: sizes the host builder would take seconds for.  Size-independent properties:
    every edge lands in exactly one party's CSR, rows are sorted, in-degrees add up, gather of ones = row lengths."""
    import torch

    import bench

    n, E, T = 200_000, 3_000_000, 4
    src, dst = bench.rmat_edges(torch, n, E, 5, "cuda")
    edges = torch.stack([src, dst], dim=1).contiguous()
    tid = (torch.arange(n, device="cuda") % T).long()
    total_edges, total_in = 0, 0
    for me in range(T):
        g = cgb.party_graph_build(edges, tid, T, me)
        total_edges += g["n_out_edges"]
        total_in += int(g["in_deg_raw"].sum())
        rp, col = g["rowptr"].long(), g["col"].long()
        assert int(rp[-1]) == g["n_out_edges"] and bool((rp[1:] >= rp[:-1]).all())
        rows = torch.repeat_interleave(torch.arange(g["n_rows"], device="cuda"), rp[1:] - rp[:-1])
        key = rows * n + col
        assert bool((key[1:] >= key[:-1]).all()), "rows ascending, sources ascending inside a row"
        assert int((src % T == me).sum()) == g["n_out_edges"]
        assert bool((g["vids"] % T == me).all()) and g["n_local"] == n // T
        ones = torch.ones((g["n_local"], 2), dtype=torch.int64, device="cuda")
        y = cgb.gather_sum(g["csr"], ones)
        assert torch.equal(y[:, 0], rp[1:] - rp[:-1])
        # dummy rule: +1 exactly for the vertices without a local in-edge
        lo, hi = g["offsets"][me], g["offsets"][me + 1]
        local_in = (rp[lo + 1:hi + 1] - rp[lo:hi])
        assert torch.equal(g["in_deg"] - g["in_deg_raw"], (local_in == 0).long())
        g["csr"].destroy()
    assert total_edges == E and total_in == E
