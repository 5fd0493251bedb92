"""CPU ORACLE (test infrastructure) for the graph tile loader and the index-vector construction.

numpy restatement of
  - graph tile loading:   /root/reference/include/graph_io_util.h:40-208, include/graph.h:607-641
  - index vectors:        /root/reference/include/ss_vertex_centric_algo_kernel.h:295-534
used to check the C++ builder in cognn_b200/host and to feed the epoch oracle.  Pinned by the worked example of
SURVEY.md section 3.6 (tests/test_graph_index.py); otherwise parity unpinned (the reference has no tests).
"""
import numpy as np


def next_pow2(n):
    """get_next_power_of_2 is absent from the reference tree; frozen as: smallest power of two >= max(n, 1)."""
    p = 1
    while p < n:
        p *= 2
    return p


class Tile:
    """What one party holds after graphTilesFromEdgeList(..., tileIndex=me) with finalize=true."""

    def __init__(self, edges, tid_of, T, me):
        edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        tid_of = np.asarray(tid_of, dtype=np.int64)
        self.T, self.me = T, me
        self.tid_of = tid_of
        self.local_vids = np.nonzero(tid_of == me)[0].astype(np.int64)  # ascending (ssk.h:462-464 sorts ids)
        src_t, dst_t = tid_of[edges[:, 0]], tid_of[edges[:, 1]]
        mine = edges[src_t == me]
        # edges are kept sorted by (src, dst) (graph.h:473-476, 636-641); repeated edges are accepted (graph.h:621)
        order = np.lexsort((mine[:, 1], mine[:, 0]))
        self.edges = mine[order]
        n = tid_of.size
        self.out_deg = np.bincount(self.edges[:, 0], minlength=n).astype(np.int64)
        # in-degree counts local in-edges (graph.h:627-632) and remote ones (graph_io_util.h:170-175)
        self.in_deg = np.bincount(edges[dst_t == me][:, 1], minlength=n).astype(np.int64)


def build_index_vectors(tile, no_dummy_edge=True):
    """ssk.h:295-504.  Returns a dict of numpy arrays / lists indexed by destination party."""
    T, me, tid_of = tile.T, tile.me, tile.tid_of
    in_deg, out_deg = tile.in_deg.copy(), tile.out_deg.copy()
    e_src, e_dst = tile.edges[:, 0], tile.edges[:, 1]
    e_dtid = tid_of[e_dst]

    # per destination (local vertex or mirror vertex): sources in edge order (ssk.h:295-314)
    src_lists = {}  # dst vid -> [list of src, list of isDummy]
    for s, d in zip(e_src.tolist(), e_dst.tolist()):
        src_lists.setdefault(d, [[], []])
        src_lists[d][0].append(s)
        src_lists[d][1].append(False)
    local = tile.local_vids.tolist()
    if not no_dummy_edge:
        # ssk.h:358-398: pad every destination's source list to a power of two
        for v in local:
            lst = src_lists.setdefault(v, [[], []])
            pad = next_pow2(len(lst[0])) - len(lst[0])
            lst[0] += [v] * pad
            lst[1] += [True] * pad
        for d, lst in src_lists.items():
            if tid_of[d] == me:
                continue
            pad = next_pow2(len(lst[0])) - len(lst[0])
            lst[0] += [lst[0][0]] * pad
            lst[1] += [True] * pad
    else:
        # ssk.h:399-436: one dummy self-edge for a local vertex without local in-edge; BOTH degrees incremented
        for v in local:
            lst = src_lists.setdefault(v, [[], []])
            if len(lst[0]) == 0:
                lst[0].append(v)
                lst[1].append(True)
                in_deg[v] += 1
                out_deg[v] += 1

    out = {
        "T": T, "me": me,
        "localVertexPos": np.array(local, dtype=np.uint64),
        "localVertexInDeg": in_deg[local].astype(np.uint64),
        "localVertexOutDeg": out_deg[local].astype(np.uint64),  # not in GraphSummary; PreScatter is handed inDeg (ssk.h:739)
        "updateSrcVertexPos": [None] * T, "updateDstVertexPos": [None] * T,
        "updateSrcOutDeg": [None] * T, "updateDstInDeg": [None] * T,
        "isUpdateSrcVertexDummy": [None] * T, "isGatherDstVertexDummy": [None] * T,
        "mirrorVertexPos": [None] * T,
    }
    for t in range(T):
        if t == me:
            ids = local
        else:
            ids = sorted(d for d in src_lists if tid_of[d] == t)
        usp, udp, uso, udi, dum = [], [], [], [], []
        for d in ids:
            srcs, isd = src_lists[d]
            usp += srcs
            udp += [d] * len(srcs)
            uso += [int(out_deg[s]) for s in srcs]
            udi += [int(in_deg[d]) if t == me else 0] * len(srcs)
            dum += isd
        out["updateSrcVertexPos"][t] = np.array(usp, dtype=np.uint64)
        out["updateDstVertexPos"][t] = np.array(udp, dtype=np.uint64)
        out["updateSrcOutDeg"][t] = np.array(uso, dtype=np.uint64)
        out["updateDstInDeg"][t] = np.array(udi, dtype=np.uint64)
        out["isUpdateSrcVertexDummy"][t] = np.array(dum, dtype=bool)
        if t == me:
            out["isGatherDstVertexDummy"][t] = np.array([src_lists[d][1][0] for d in ids], dtype=bool)
        else:
            out["mirrorVertexPos"][t] = np.array(ids, dtype=np.uint64)
    out["_in_deg_after"] = in_deg
    return out


def exchange_pos_vectors(per_party):
    """ssk.h:506-534: party i sends updateDstVertexPos[j] to j; j derives remoteMirrorVertexPos[i],
    remoteUpdateDstInDeg[i] and isGatherDstVertexDummy[i]."""
    T = len(per_party)
    for me, iv in enumerate(per_party):
        iv["remoteMirrorVertexPos"] = [None] * T
        iv["remoteUpdateDstInDeg"] = [None] * T
        pos_index = {int(v): k for k, v in enumerate(iv["localVertexPos"].tolist())}
        in_deg = iv["_in_deg_after"]
        for i in range(T):
            if i == me:
                continue
            rm = per_party[i]["updateDstVertexPos"][me]
            iv["remoteMirrorVertexPos"][i] = rm.copy()
            iv["remoteUpdateDstInDeg"][i] = np.array([int(in_deg[int(v)]) for v in rm.tolist()], dtype=np.uint64)
            dummy = np.ones(iv["localVertexPos"].size, dtype=bool)
            for v in rm.tolist():
                dummy[pos_index[int(v)]] = False
            iv["isGatherDstVertexDummy"][i] = dummy
    return per_party


def build_all(edges, tid_of, T, no_dummy_edge=True):
    tiles = [Tile(edges, tid_of, T, p) for p in range(T)]
    ivs = [build_index_vectors(t, no_dummy_edge) for t in tiles]
    return tiles, exchange_pos_vectors(ivs)


def csr_from_pos(src_pos, dst_pos, src_ids, dst_ids, drop=None):
    """CSR-by-destination over row indices: rows = position of dst in dst_ids, cols = position of src in src_ids.
    dst_pos must be grouped by destination ascending (it is, ssk.h:462-504).  `drop`: bool mask of entries to skip
    (dummy edges) when building the fused operator."""
    src_index = {int(v): k for k, v in enumerate(np.asarray(src_ids).tolist())}
    dst_index = {int(v): k for k, v in enumerate(np.asarray(dst_ids).tolist())}
    n_rows = len(dst_index)
    counts = np.zeros(n_rows, dtype=np.int64)
    cols = []
    for k, (s, d) in enumerate(zip(np.asarray(src_pos).tolist(), np.asarray(dst_pos).tolist())):
        if drop is not None and drop[k]:
            continue
        counts[dst_index[int(d)]] += 1
        cols.append(src_index[int(s)])
    rowptr = np.zeros(n_rows + 1, dtype=np.uint32)
    rowptr[1:] = np.cumsum(counts)
    return rowptr, np.array(cols, dtype=np.uint32)
