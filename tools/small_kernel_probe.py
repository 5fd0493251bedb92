#!/usr/bin/env python
"""Steady-state time per launch of the kernels one secure-GCN epoch issues, at the shapes of a small dataset (default: Cora
split over 2 parties: 1354 vertices per party, F = 1433, H = 16, C = 7, ~5.3k edges per party CSR).

Each kernel is launched `--reps` times back to back on one stream between two CUDA events, after a warm-up: what a CUDA-graph
replay pays per node once clocks are up and the operands sit in L2 (an ncu launch list isolates every kernel and reports
2-5x more for these latency-bound launches).  Used to pick the chunk size of small CSRs and the split-K depth of the
integer-pipe matmul.  Prints one JSON line.

  python tools/small_kernel_probe.py [--n 1354 --F 1433 --H 16 --C 7 --edges 5300]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cognn_b200 as cg  # noqa: E402


STREAM = None


def timed(fn, reps):
    """`reps` launches recorded into one CUDA graph (ctypes costs more per call than these kernels run), replayed 5 times."""
    with torch.cuda.stream(STREAM):
        for _ in range(5):
            fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=STREAM):
        for _ in range(reps):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (5 * reps)  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1354)
    ap.add_argument("--F", type=int, default=1433)
    ap.add_argument("--H", type=int, default=16)
    ap.add_argument("--C", type=int, default=7)
    ap.add_argument("--edges", type=int, default=5300)
    ap.add_argument("--rows", type=int, default=2708)
    ap.add_argument("--reps", type=int, default=100)
    args = ap.parse_args()
    global STREAM
    dev = torch.device("cuda", 0)
    STREAM = torch.cuda.Stream(dev)
    with torch.cuda.stream(STREAM):
        ctx = cg.Context(0)  # bound to STREAM, the stream the graphs are captured on
    g = torch.Generator(device="cpu").manual_seed(1)

    def rnd(*shape):
        return torch.randint(-2**62, 2**62, shape, dtype=torch.int64, generator=g).to(dev)

    n, F, H, C = args.n, args.F, args.H, args.C
    out = {"shape": {"n": n, "F": F, "H": H, "C": C, "edges": args.edges, "rows": args.rows},
           "env": {k: v for k, v in os.environ.items() if k.startswith("CGB_")}, "unit": "us per launch"}
    # the party CSR: `rows` destination rows (own + mirrors), `edges` edges with sources among the party's n vertices
    rs = np.random.default_rng(3)
    dst = np.sort(rs.integers(0, args.rows, size=args.edges))
    src = rs.integers(0, n, size=args.edges).astype(np.int32)
    rowptr = np.zeros(args.rows + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    csr = ctx.csr_create(torch.from_numpy(rowptr), torch.from_numpy(src), n)
    for D in (H, C):
        x, delta, y = rnd(n, D), rnd(args.rows, D), ctx.empty(args.rows, D)
        out[f"gather_sum D={D}"] = round(timed(lambda: ctx.gather_sum(csr, x, delta, out=y), args.reps), 2)
    # the five Beaver products of an epoch (M, K, N)
    for name, (M, K, N) in {"X*W0": (n, F, H), "H*W1": (n, H, C), "g*W1^T": (n, C, H), "h1^T*v": (H, n, C),
                            "h0^T*v": (F, n, H)}.items():
        U, V, Z = rnd(M, K), rnd(K, N), rnd(M, N)
        mine, peer = rnd(M * K + K * N), rnd(M * K + K * N)
        for share in (0, 1):
            out[f"matmul_finish_open {name} {M}x{K}x{N} share{share}"] = round(
                timed(lambda: ctx.beaver_matmul_finish_open(mine, peer, U, V, Z, share), args.reps), 2)
    key = [1, 2, 3, 4, 5, 6, 7, 8]
    for D in (H, C):
        a, b, o = rnd(n * D), rnd(n * D), ctx.empty(n * D)
        out[f"prg_sum D={D} (2 in, 1 stream)"] = round(timed(lambda: ctx.prg_sum(key, [5], [a, b], out=o), args.reps), 2)
        out[f"prg_mask_sub D={D}"] = round(timed(lambda: ctx.prg_mask_sub(key, 5, 0, a, out=o), args.reps), 2)
        out[f"add D={D}"] = round(timed(lambda: ctx.add(a, b, out=o), args.reps), 2)
    xf = rnd(n, F)
    out["transpose n x F"] = round(timed(lambda: ctx.transpose(xf), args.reps), 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
