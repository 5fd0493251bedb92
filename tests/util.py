"""Shared helpers for the parity tests: seeded inputs, numpy <-> torch (u64 bits in int64 tensors)."""
import numpy as np


def rand_u64(rng, *shape):
    return rng.integers(0, 1 << 64, size=shape, dtype=np.uint64)


def to_dev(a, device="cuda"):
    import torch

    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        return torch.from_numpy(a.view(np.int64)).to(device)
    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32)).to(device)
    return torch.from_numpy(a).to(device)


def to_np(t):
    a = t.detach().cpu().numpy()
    if a.dtype == np.int64:
        return a.view(np.uint64)
    if a.dtype == np.int32:
        return a.view(np.uint32)
    return a


def power_law_csr(rng, n_dst, n_src, n_edges, max_deg=None, zipf_a=1.6):
    """Destination degrees ~ Zipf (power law), a few empty rows, sources skewed too.  Returns (rowptr, col) uint32."""
    w = rng.zipf(zipf_a, size=n_dst).astype(np.float64)
    if max_deg is not None:
        w = np.minimum(w, max_deg)
    w[rng.integers(0, n_dst, size=max(1, n_dst // 50))] = 0  # empty rows
    deg = np.floor(w / max(w.sum(), 1) * n_edges).astype(np.int64)
    rowptr = np.zeros(n_dst + 1, dtype=np.uint32)
    rowptr[1:] = np.cumsum(deg)
    e = int(rowptr[-1])
    col = np.minimum((rng.pareto(1.2, size=e) * n_src / 50).astype(np.int64), n_src - 1)
    perm = rng.permutation(n_src)
    return rowptr, perm[col].astype(np.uint32)
