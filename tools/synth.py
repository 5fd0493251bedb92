"""Synthetic inputs of the named shapes: symmetric directed edge lists (each undirected pair stored
twice, like tools/data_transform.py of the reference), Chung-Lu power-law degrees (exponent 2.1), no self loops,
exactly E directed entries; features / labels / partition as the survey specifies.  Seeds: graph 42, features 43,
labels 44."""
import numpy as np

SHAPES = {
    # name: (N, E directed entries, F, H, C, feature kind)
    "cora_small": (4, 8, 2, 3, 3, "dense"),
    "cora": (2708, 10556, 1433, 16, 7, "bow"),
    "citeseer": (3327, 9104, 3703, 16, 6, "bow"),
    "pubmed": (19717, 88648, 500, 16, 3, "dense"),
    "arxiv": (169343, 1166243 - 1166243 % 2, 128, 16, 40, "dense"),
    "arxiv256": (169343, 1166243 - 1166243 % 2, 128, 256, 40, "dense"),
}


def chung_lu_pairs(n, n_pairs, rng, exponent=2.1, allowed=None):
    """n_pairs distinct undirected pairs (a < b) with P(a, b) ~ w_a w_b, w_i ~ (i+1)^(-1/(exponent-1))."""
    w = (np.arange(n) + 1.0) ** (-1.0 / (exponent - 1.0))
    w = rng.permutation(w)
    cdf = np.cumsum(w) / w.sum()
    got = np.zeros(0, dtype=np.int64)
    while got.size < n_pairs:
        k = int((n_pairs - got.size) * 1.3) + 16
        a = np.searchsorted(cdf, rng.random(k))
        b = np.searchsorted(cdf, rng.random(k))
        a, b = np.minimum(a, n - 1), np.minimum(b, n - 1)
        keep = a != b
        if allowed is not None:
            keep &= allowed(a, b)
        lo, hi = np.minimum(a, b)[keep], np.maximum(a, b)[keep]
        got = np.unique(np.concatenate([got, lo * n + hi]))
    got = rng.permutation(got)[:n_pairs]
    return got // n, got % n


def make(shape, T, inter_fraction=None):
    N, E, F, H, C, kind = SHAPES[shape]
    rng = np.random.default_rng(42)
    if shape == "cora_small":
        a, b = np.array([0, 1, 2, 0]), np.array([1, 2, 3, 2])
        tid = np.arange(N) % T
    elif inter_fraction is None:
        a, b = chung_lu_pairs(N, E // 2, rng)
        tid = np.arange(N) % T  # tools/data_transform.py:25
    else:
        # BASELINE configs[2]: contiguous blocks, exactly `inter_fraction` of the edges cross parties
        tid = np.minimum(np.arange(N) * T // N, T - 1)
        n_inter = int(round(E // 2 * inter_fraction))
        ai, bi = chung_lu_pairs(N, n_inter, rng, allowed=lambda x, y: tid[x] != tid[y])
        al, bl = chung_lu_pairs(N, E // 2 - n_inter, rng, allowed=lambda x, y: tid[x] == tid[y])
        a, b = np.concatenate([ai, al]), np.concatenate([bi, bl])
    edges = np.stack([np.concatenate([a, b]), np.concatenate([b, a])], axis=1).astype(np.int64)
    frng = np.random.default_rng(43)
    if kind == "bow":
        feats = (frng.random((N, F)) < 0.0125).astype(np.float64)
    else:
        feats = frng.random((N, F))
    labels = np.random.default_rng(44).integers(0, C, size=N).astype(np.int32)
    cfg = dict(input_dim=F, hidden_dim=H, num_labels=C, num_samples=N, num_edges=int(edges.shape[0]),
               learning_rate=0.5, train_ratio=0.2, val_ratio=0.2, test_ratio=0.6)
    inter = int((tid[edges[:, 0]] != tid[edges[:, 1]]).sum())
    return {"edges": edges, "tid": tid.astype(np.int64), "feats": feats, "labels": labels, "cfg": cfg,
            "N": N, "E": int(edges.shape[0]), "inter_party_edges": inter}


def write_reference_files(g, prefix):
    """The reference's three text formats (graph_io_util.h:66-147, kernel_harness.h:37-44, data_transform.py:19-62)."""
    with open(prefix + ".edge.preprocessed", "w") as f:
        for s, d in g["edges"].tolist():
            f.write(f"{s} {d}\n")
    with open(prefix + ".part.preprocessed", "w") as f:
        for v, t in enumerate(g["tid"].tolist()):
            f.write(f"{v} {t}\n")
    with open(prefix + ".vertex.preprocessed", "w") as f:
        for v in range(g["N"]):
            f.write(f"{v} " + " ".join(repr(float(x)) for x in g["feats"][v]) + f" {int(g['labels'][v])}\n")
