"""Seeded synthetic graphs shared by the epoch tests (symmetric edge lists, like tools/data_transform.py output)."""
import numpy as np


def small_graph(n, n_edges, F, C, T, seed, partition="mod", multi_edges=0, isolated=0):
    """multi_edges: how many existing directed entries to repeat (repeated edges are accepted, graph.h:621);
    isolated: how many trailing vertices get no edge at all (zero in-degree: dummy self-edge rule, ssk.h:412-418)."""
    rng = np.random.default_rng(seed)
    n_full, n = n, n - isolated
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
    for k in range(multi_edges):
        edges.append(edges[int(rng.integers(0, len(edges)))])
    n = n_full
    tid = (np.arange(n) % T) if partition == "mod" else (np.arange(n) * T // n)
    feats = (rng.random((n, F)) < 0.3).astype(np.float64) * rng.random((n, F))
    labels = rng.integers(0, C, size=n)
    return {"edges": np.array(edges, dtype=np.int64), "tid": tid.astype(np.int64), "feats": feats, "labels": labels}
