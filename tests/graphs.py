"""Seeded synthetic graphs shared by the epoch tests (symmetric edge lists, like tools/data_transform.py output)."""
import numpy as np


def small_graph(n, n_edges, F, C, T, seed, partition="mod"):
    rng = np.random.default_rng(seed)
    pairs = set()
    # skewed endpoints so a few vertices get many neighbours; guarantee some isolated / remote-only vertices
    while len(pairs) < n_edges // 2:
        a = int(min(n - 1, rng.pareto(1.5) * n / 10))
        b = int(rng.integers(0, n))
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    edges = []
    for a, b in sorted(pairs):
        edges += [(a, b), (b, a)]
    tid = (np.arange(n) % T) if partition == "mod" else (np.arange(n) * T // n)
    feats = (rng.random((n, F)) < 0.3).astype(np.float64) * rng.random((n, F))
    labels = rng.integers(0, C, size=n)
    return {"edges": np.array(edges, dtype=np.int64), "tid": tid.astype(np.int64), "feats": feats, "labels": labels}
