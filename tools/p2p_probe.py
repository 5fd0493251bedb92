"""NVLink transport probe for the mirror-update exchange (ssk.h:835 -> 1067/1090): how fast can one n_local x D block per
ordered pair of parties move between GPUs, with which engine, and what does it cost the gather kernel that runs beside it?

Run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_probe.py
Prints one JSON line per experiment on rank 0 (times are the max over ranks of CUDA-event durations).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=6_250_000)
    ap.add_argument("--dim", type=int, default=16)
    ap.add_argument("--edges", type=int, default=100_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-gather", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import cognn_b200
    from bench import RawCuda, build_party_csr

    rank, P, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ctx = cognn_b200.Context(lr)
    n, D = args.rows, args.dim
    blk = n * D * 8
    # two exported buffers per rank: STAGE (what this rank produced, one block per destination) and WIN (what it receives)
    stage_p, win_p = ctx.malloc(P * blk), ctx.malloc(P * blk)
    everyone = [None] * P
    dist.all_gather_object(everyone, (ctx.ipc_export(stage_p), ctx.ipc_export(win_p)))
    peer_stage = [stage_p if t == rank else ctx.ipc_open(everyone[t][0]) for t in range(P)]
    peer_win = [win_p if t == rank else ctx.ipc_open(everyone[t][1]) for t in range(P)]
    stage = torch.as_tensor(RawCuda(stage_p, (P, n, D)), device=dev)
    win = torch.as_tensor(RawCuda(win_p, (P, n, D)), device=dev)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    stage.copy_(torch.randint(-2**62, 2**62, (P, n, D), dtype=torch.int64, device=dev, generator=g))
    win.zero_()
    streams = [torch.cuda.Stream(device=dev) for _ in range(P)]
    sctx = []
    for t in range(P):
        with torch.cuda.stream(streams[t]):
            sctx.append(cognn_b200.Context(lr))
    main_s = torch.cuda.current_stream()
    peers = [(rank + j) % P for j in range(1, P)]
    flag = torch.zeros(1, dtype=torch.int32, device=dev)

    def view(ptr, shape):
        return torch.as_tensor(RawCuda(ptr, shape), device=dev)

    def timed(name, fn, nbytes_out, extra=None):
        """fn() enqueues the work of one repetition on any streams and returns the list of streams to join."""
        best = None
        for _ in range(args.reps):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main_s)
            for s in streams:
                s.wait_event(e0)
            used = fn() or []
            for s in used:
                ev = torch.cuda.Event()
                ev.record(s)
                main_s.wait_event(ev)
            e1.record(main_s)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            best = ms if best is None else min(best, ms)
        if rank == 0:
            line = {"exp": name, "n_gpus": P, "ms": round(best, 4), "egress_GBps_per_gpu": round(nbytes_out / best / 1e6, 1)}
            if extra:
                line.update(extra())
            print(json.dumps(line), flush=True)
        return best

    # ---- copy engines -------------------------------------------------------------------------------------------
    def ce(targets):
        def fn():
            for t in targets:
                with torch.cuda.stream(streams[t]):
                    view(peer_win[t] + rank * blk, (n, D)).copy_(stage[t], non_blocking=True)
            return [streams[t] for t in targets]
        return fn

    timed("ce_push_ring", ce(peers[:1]), blk)
    if P > 2:
        timed("ce_push_all", ce(peers), blk * (P - 1))

    # ---- SM copy kernel, push and pull ------------------------------------------------------------------------------
    def sm(targets, n_ctas, pull):
        def fn():
            for t in targets:
                if pull:  # read block `rank` of peer t's stage into my window slot t
                    sctx[t].peer_copy(win_p + t * blk, peer_stage[t] + rank * blk, blk, n_ctas)
                else:
                    sctx[t].peer_copy(peer_win[t] + rank * blk, stage_p + t * blk, blk, n_ctas)
            return [streams[t] for t in targets]
        return fn

    for n_ctas in (8, 16, 32, 64, 148, 296):
        timed(f"sm_push_ring_ctas{n_ctas}", sm(peers[:1], n_ctas, False), blk)
    for n_ctas in (16, 32, 64, 148):
        timed(f"sm_pull_ring_ctas{n_ctas}", sm(peers[:1], n_ctas, True), blk)
    if P > 2:
        for n_ctas in (8, 16, 32, 64):
            timed(f"sm_push_all_ctas{n_ctas}_per_peer", sm(peers, n_ctas, False), blk * (P - 1))
        for n_ctas in (8, 16, 32, 64):
            timed(f"sm_pull_all_ctas{n_ctas}_per_peer", sm(peers, n_ctas, True), blk * (P - 1))

    # ---- pull fused into the GatherComp sum: V = sum over source parties, reading the peers' stage blocks directly -------
    v = torch.empty((n, D), dtype=torch.int64, device=dev)
    srcs = [view(peer_stage[t] + rank * blk, (n, D)) for t in range(P)]

    def fused_sum():
        ctx.sum_n(srcs, out=v)
        return []

    timed("sum_n_reading_peer_blocks", fused_sum, blk * (P - 1))
    ref = stage[rank].clone()
    dist.barrier()
    # check: same as NCCL all-to-all + local sum
    recv = torch.empty_like(stage)
    dist.all_to_all_single(recv.view(P * n, D), stage.view(P * n, D))
    ref = recv.sum(0)
    assert torch.equal(ref, v), "peer-read sum differs from all-to-all + sum"

    def nccl_a2a():
        dist.all_to_all_single(recv.view(P * n, D), stage.view(P * n, D))
        return []

    timed("nccl_all_to_all", nccl_a2a, blk * (P - 1))

    # ---- beside a gather kernel ---------------------------------------------------------------------------------------
    if not args.no_gather:
        rowptr, col = build_party_csr(torch, n, args.edges, 1, rank, 42, dev)
        csr = ctx.csr_create(rowptr, col, n)
        x = stage[rank]
        y = torch.empty((n, D), dtype=torch.int64, device=dev)
        gt = {}

        def with_gather(name, side, nb=blk):
            ga, gb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

            def fn():
                used = side() if side else []
                ga.record(main_s)
                for _ in range(2):
                    ctx.gather_sum(csr, x, None, out=y)
                gb.record(main_s)
                return used

            def extra():
                return {"gather_ms_per_launch_rank0": round(ga.elapsed_time(gb) / 2, 4)}

            gt[name] = timed(name, fn, nb if side else 0, extra)

        with_gather("gather_x2_alone", None)
        with_gather("gather_x2_with_ce_push_ring", ce(peers[:1]))
        for n_ctas in (16, 32, 64):
            with_gather(f"gather_x2_with_sm_push_ring_ctas{n_ctas}", sm(peers[:1], n_ctas, False))
            with_gather(f"gather_x2_with_sm_pull_ring_ctas{n_ctas}", sm(peers[:1], n_ctas, True))
        if P > 2:
            with_gather("gather_x2_with_ce_push_all", ce(peers), blk * (P - 1))
            with_gather("gather_x2_with_sm_push_all_ctas16", sm(peers, 16, False), blk * (P - 1))
            with_gather("gather_x2_with_sm_pull_all_ctas16", sm(peers, 16, True), blk * (P - 1))

    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
