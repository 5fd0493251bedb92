#!/usr/bin/env python
"""What bounds the D = 16 gather?  Three measurements with the product kernels, one JSON line each:

  stream_l2   cgb_sum_n over 8 aliases of one 32 MB block (stays in the 126 MB L2): bytes the L2 slices deliver per second to a
              perfectly coalesced streaming reader -- the practical L2 -> SM ceiling of this device;
  gather_l2   the gather kernel on a graph whose share rows fit in L2 (uniform random sources over 2^18 rows = 32 MB, E = 2^26):
              every row read is an L2 hit, no DRAM fill traffic -- the ceiling of the kernel's own access pattern (random 128-byte
              rows, one LDG.E.256 per four lanes);
  gather_dram the same kernel, uniform random sources over 6.25M rows (800 MB): every row read is a DRAM read.

The power-law bench graph sits between the last two (58 % of its sectors hit L2).  ncu of the bench launch
(profiles/r1b_ncu_gather_chunk_vec4_chunk128.txt) shows lts__t_sectors = 636M sectors in 1.766 ms = 11.5 TB/s of L2 slice traffic for
14.0 GB of algorithmic row bytes: compare with stream_l2 / gather_l2 here."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(torch, fn, reps):
    fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    import torch

    import bench
    import cognn_b200

    dev = torch.device("cuda", 0)
    ctx = cognn_b200.Context(0)
    g = torch.Generator(device=dev).manual_seed(1)
    # 1. streaming reads out of L2
    n = 32 * 1024 * 1024 // 8
    blk = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device=dev, generator=g)
    out = torch.empty_like(blk)
    ms = timed(torch, lambda: ctx.sum_n([blk] * 8, out=out), 20)
    print(json.dumps({"probe": "stream_l2", "bytes_read": 8 * n * 8, "bytes_written": n * 8, "ms": ms,
                      "l2_read_GBps": 8 * n * 8 / ms / 1e6}), flush=True)
    del blk, out
    D = 16
    for name, n_src, E in (("gather_l2", 1 << 18, 1 << 26), ("gather_dram", 6_250_000, 100_000_000)):
        src = torch.randint(0, n_src, (E,), dtype=torch.int64, device=dev, generator=g)
        n_dst = E // 16
        dst = torch.randint(0, n_dst, (E,), dtype=torch.int64, device=dev, generator=g)
        rowptr, col = bench.csr_from_edges(torch, src, dst, n_src, n_dst)
        del src, dst
        csr = ctx.csr_create(rowptr, col, n_src)
        x = torch.randint(-2**63, 2**63 - 1, (n_src, D), dtype=torch.int64, device=dev, generator=g)
        y = torch.empty((n_dst, D), dtype=torch.int64, device=dev)
        ms = timed(torch, lambda: ctx.gather_sum(csr, x, None, out=y), 10)
        alg = bench.algorithmic_bytes(n_dst, E, D)
        print(json.dumps({"probe": name, "n_src_rows": n_src, "x_MB": n_src * D * 8 / 1e6, "edges": E, "out_rows": n_dst, "ms": ms,
                          "edges_per_s": E / ms * 1e3, "row_bytes_GBps": E * D * 8 / ms / 1e6, "algorithmic_GBps": alg / ms / 1e6,
                          "kernel": ctx.last_kernel}), flush=True)
        csr.destroy()
        del rowptr, col, x, y
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
