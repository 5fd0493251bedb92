#!/usr/bin/env python
"""BASELINE.json configs[4]: share-local kernel sweep on one GPU (gather-sum over RMAT graphs of 1M-100M edges,
u64 matmul N x F . F x H with F, H in {128, 256, 512}) plus stream kernels.  Every timed launch is preceded by an
L2 flush (256 MB write) unless the working set is far above L2; times are CUDA events on the launch stream.
Writes one JSON object per line to stdout (and --out)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import cognn_b200  # noqa: E402


def timed(fn, reps, flush):
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--gather", default="1000000,10000000,100000000")
    ap.add_argument("--dims", default="16,64,128")
    ap.add_argument("--matmul-rows", type=int, default=1 << 20)
    ap.add_argument("--skip-matmul", action="store_true")
    ap.add_argument("--mm", default=None, help="only this matmul shape, e.g. 512x512 (K x N)")
    ap.add_argument("--skip-stream", action="store_true")
    ap.add_argument("--uniform", action="store_true", help="add a uniform-degree control graph at the largest size")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = cognn_b200.Context(0)
    peak, _ = bench.measured_peak_hbm()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []

    def emit(rec):
        out.append(rec)
        print(json.dumps(rec), flush=True)

    g = torch.Generator(device=dev).manual_seed(43)
    for E in [int(v) for v in args.gather.split(",") if v]:
        n = max(1, E // 16)
        graphs = [("rmat", bench.build_party_csr(torch, n, E, 1, 0, 42, dev))]
        if args.uniform and E == max(int(v) for v in args.gather.split(",")):
            col = torch.randint(0, n, (n * 16,), device=dev, generator=g, dtype=torch.int64).int()
            rowptr = (torch.arange(n + 1, device=dev, dtype=torch.int64) * 16).int()
            graphs.append(("uniform16", (rowptr, col)))
        for name, (rowptr, col) in graphs:
            csr = ctx.csr_create(rowptr, col, n)
            deg = (rowptr[1:] - rowptr[:-1])
            for D in [int(v) for v in args.dims.split(",") if v]:
                x = torch.randint(-2**63, 2**63 - 1, (n, D), dtype=torch.int64, device=dev, generator=g)
                y = torch.empty((n, D), dtype=torch.int64, device=dev)
                for _ in range(3):
                    ctx.gather_sum(csr, x, None, out=y)
                med, best = timed(lambda: ctx.gather_sum(csr, x, None, out=y), 7, flush)
                alg = bench.algorithmic_bytes(n, col.numel(), D)
                emit({"kernel": "gather_sum", "graph": name, "edges": int(col.numel()), "rows": n, "D": D,
                      "max_deg": int(deg.max()), "ms_median": med, "ms_best": best,
                      "edges_per_s": col.numel() / (med * 1e-3), "alg_GBps": alg / (med * 1e-3) / 1e9,
                      "frac_of_measured_hbm": alg / (med * 1e-3) / 1e9 / peak,
                      "frac_of_8TBps": alg / (med * 1e-3) / 1e9 / 8000.0,
                      "impl": os.environ.get("CGB_GATHER_IMPL", "chunks"), "u4": bool(os.environ.get("CGB_GATHER_U4"))})
                del x, y
            csr.destroy()
            del rowptr, col

    if not args.skip_matmul:
        M = args.matmul_rows
        Fs, Hs = (128, 256, 512), (16, 128, 256, 512)
        if args.mm:
            Fs, Hs = (int(args.mm.split("x")[0]),), (int(args.mm.split("x")[1]),)
        for F in Fs:
            A = torch.randint(-2**63, 2**63 - 1, (M, F), dtype=torch.int64, device=dev, generator=g)
            for H in Hs:
                B = torch.randint(-2**63, 2**63 - 1, (F, H), dtype=torch.int64, device=dev, generator=g)
                C = torch.empty((M, H), dtype=torch.int64, device=dev)
                for _ in range(2):
                    ctx.matmul(A, B, out=C)
                med, best = timed(lambda: ctx.matmul(A, B, out=C), 5, None)
                macs = M * F * H
                emit({"kernel": "matmul_u64", "impl": os.environ.get("CGB_MATMUL_IMPL", "auto"), "M": M, "K": F, "N": H, "ms_median": med, "ms_best": best,
                      "u64_mac_per_s": macs / (med * 1e-3), "bytes_GBps": 8 * (M * F + F * H + M * H) / (med * 1e-3) / 1e9})
                del B, C
            del A
        # weight-gradient shape (split-K): X^T (F x N_p) * G (N_p x H)
        for (Np, F, H) in ([] if args.mm else [(21168, 128, 256), (1 << 20, 128, 16)]):
            X = torch.randint(-2**63, 2**63 - 1, (Np, F), dtype=torch.int64, device=dev, generator=g)
            G = torch.randint(-2**63, 2**63 - 1, (Np, H), dtype=torch.int64, device=dev, generator=g)
            for _ in range(2):
                ctx.matmul(X, G, transA=True)
            med, best = timed(lambda: ctx.matmul(X, G, transA=True), 5, flush)
            emit({"kernel": "matmul_u64_transA_splitK", "M": F, "K": Np, "N": H, "ms_median": med, "ms_best": best,
                  "u64_mac_per_s": Np * F * H / (med * 1e-3)})

    if args.skip_stream:
        n = 0
    else:
        n = 256 << 20
    if n == 0:
        if args.out:
            with open(args.out, "w") as f:
                for r in out:
                    f.write(json.dumps(r) + "\n")
        return
    # stream kernels: 256 Mi words (2 GB per buffer)
    a = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device=dev, generator=g)
    b = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device=dev, generator=g)
    o = torch.empty_like(a)
    key = [45, 0, 0, 0, 0, 0, 0, 1]
    for name, fn, byts in [
        ("ew_add", lambda: ctx.add(a, b, out=o), 24 * n),
        ("ew_scale_public_trunc", lambda: ctx.scale_public(a, 12345, 1, 16, out=o), 16 * n),
        ("prg_fill", lambda: ctx.prg_fill(key, 1, 0, n, out=o), 8 * n),
        ("prg_mask_sub", lambda: ctx.prg_mask_sub(key, 1, 0, a, out=o), 16 * n),
    ]:
        for _ in range(2):
            fn()
        med, best = timed(fn, 5, None)
        emit({"kernel": name, "words": n, "ms_median": med, "GBps": byts / (med * 1e-3) / 1e9,
              "frac_of_measured_hbm": byts / (med * 1e-3) / 1e9 / peak})
    if args.out:
        with open(args.out, "w") as f:
            for r in out:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
