#!/usr/bin/env python
"""Small single-GPU launchers for ncu captures of kernels that bench.py only reaches with several ranks or inside records:

  python tools/prof_kernels.py signal   the fused gather + signalling launch on one party's graph of the 2-party bench workload
                                        (100M edges, own block dense, 4 remote pieces compact, flags in local memory)
  python tools/prof_kernels.py matmul   the persistent tcgen05 limb matmul, 2^20 x 512 x 512 (limb split + tensor kernel)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    import cognn_b200

    mode = sys.argv[1] if len(sys.argv) > 1 else "signal"
    dev = torch.device("cuda", 0)
    ctx = cognn_b200.Context(0)
    g = torch.Generator(device=dev).manual_seed(1)
    if mode == "signal":
        E, P, D = 100_000_000, 2, 16
        n = E // 16
        rowptr, col = bench.build_party_csr(torch, n, E, P, 0, 42, dev, own_first=True)
        csr = ctx.csr_create(rowptr, col, n)
        del rowptr, col
        S = bench.sub_blocks(P)
        cuts = [n * u // S for u in range(S + 1)]
        offsets = [0] + [n + cuts[u] for u in range(S)] + [2 * n]
        nz = csr.nonempty_rows().long()
        b = torch.searchsorted(nz, torch.tensor(offsets, device=dev)).tolist()
        x = torch.randint(-2**63, 2**63 - 1, (n, D), dtype=torch.int64, device=dev, generator=g)
        v = torch.empty((n, D), dtype=torch.int64, device=dev)
        bufs = [torch.empty((max(1, b[i + 1] - b[i]), D), dtype=torch.int64, device=dev) for i in range(1, S + 1)]
        flags = torch.zeros(S + 1, dtype=torch.int32, device=dev)
        for step in range(1, 6):
            ctx.gather_sum_signal(csr, x, offsets, [v.data_ptr()] + [t.data_ptr() for t in bufs], [False] + [True] * S,
                                  [flags.data_ptr() + 4 * i for i in range(S + 1)], step)
        torch.cuda.synchronize()
        assert flags.tolist() == [5] * (S + 1)
        print("signal ok", ctx.last_kernel)
    else:
        M, F, H = 1 << 20, 512, 512
        A = torch.randint(-2**63, 2**63 - 1, (M, F), dtype=torch.int64, device=dev, generator=g)
        B = torch.randint(-2**63, 2**63 - 1, (F, H), dtype=torch.int64, device=dev, generator=g)
        C = torch.empty((M, H), dtype=torch.int64, device=dev)
        ctx.set_matmul_impl("tc")
        for _ in range(3):
            ctx.matmul(A, B, out=C)
        torch.cuda.synchronize()
        print("matmul ok", ctx.last_kernel)


if __name__ == "__main__":
    main()
