#!/usr/bin/env python
"""bench.py -- share-gather throughput of the secret-shared GCN path (BASELINE.json metric), B200 vs host CPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference [...]                          the reference's CPU path (oracle port, see below)
  torchrun ... bench.py --gpus N ...                              one rank per GPU == one party per GPU

Workload (config.workload): BASELINE.json configs[4], the largest single-GPU configuration -- gather-sum of Z_2^64
share rows over a synthetic RMAT power-law graph, E = 100M edges, N = E/16 vertices, D = 16 u64 columns (hidden_dim
of every reference config).  One "step" = one pass of the fused scatter/gather-sum (expand -> ScatterComp copy ->
prefix_network_aggregate -> extract of one GAS iteration, ss_vertex_centric_algo_kernel.h:751-821) over all edges
a party owns.  With N > 1 ranks every rank is one party holding E edges whose destinations are spread over all
parties by `vid % N` (tools/data_transform.py:25): the step is the gather into one block per destination party, the
exchange of those mirror-update blocks (ssk.h:835 -> 1067/1090) and the share-local sum of the received
blocks into the party's vertex rows (GatherComp add, optimize-gcn/gcn.h:456).  Weak scaling: edges per GPU are fixed.
WAN / 2PC-residual costs are out of scope on both arms.

Every line checks itself: the result of the TIMED configuration is compared with the CPU oracle (`parity`), the
multi-GPU exchange with a plain gather + NCCL all-to-all + sum of the same inputs (`exchange_check`), the matmul
records with the oracle on sampled rows, and the secure-GCN epoch with the epoch oracle (`bit_exact_vs_oracle`).

The reference cannot be built here (its arithmetic is in un-vendored trees), so `--impl reference` and
`cpu_baseline` time the CPU oracle (oracle/cgb_oracle.c, a restatement of the same path; kind = "port") with all
host threads.  Under N ranks the reference arm does the same N-party work on the host: N gathers of E edges and the block sums.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "share_gather_edges_per_sec"
UNIT = "edges/s"


# ------------------------------------------------------------------------------------------------------------
# synthetic graph (torch; runs on the GPU for the full size, on the CPU for small local checks)
# ------------------------------------------------------------------------------------------------------------
def rmat_edges(torch, n_vertices, n_edges, seed, device, a=0.57, b=0.19, c=0.19):
    """RMAT(a,b,c) directed edge list folded onto [0, n_vertices): returns (src, dst) int64."""
    g = torch.Generator(device=device).manual_seed(seed)
    bits = max(1, (n_vertices - 1).bit_length())
    src = torch.zeros(n_edges, dtype=torch.int64, device=device)
    dst = torch.zeros(n_edges, dtype=torch.int64, device=device)
    for _ in range(bits):
        r = torch.rand(n_edges, device=device, generator=g)
        sbit = (r >= a + b).long()                       # quadrants c, d -> source bit 1
        dbit = ((r >= a) & (r < a + b) | (r >= a + b + c)).long()  # quadrants b, d -> destination bit 1
        src = (src << 1) | sbit
        dst = (dst << 1) | dbit
        del r, sbit, dbit
    # decorrelate ids from degree before folding (RMAT puts the hubs at small ids AND skews every id bit: 76 % of the
    # destinations are even, which a `vid % T` partition would turn into a 76/24 load split no real graph has -- Graph500
    # scrambles labels for the same reason).  A multiply-xorshift mix; a plain odd multiplier keeps the parity skew.
    def mix(v, c1, c2):
        m63 = (1 << 63) - 1
        v = (v * c1) & m63
        v = v ^ (v >> 31)
        v = (v * c2) & m63
        return v ^ (v >> 29)

    src = mix(src, 2654435761, 0x9E3779B97F4A7C15 & ((1 << 63) - 1)) % n_vertices
    dst = mix(dst + 54321, 2246822519, 0xBF58476D1CE4E5B9 & ((1 << 63) - 1)) % n_vertices
    return src, dst


def csr_from_edges(torch, src, out_row, n_src, n_rows):
    """CSR-by-destination: rows sorted, sources ascending inside a row (ssk.h:295-314 order)."""
    key = torch.sort(out_row * n_src + src).values
    out_row = key // n_src
    col = (key - out_row * n_src).int()
    del key
    counts = torch.bincount(out_row, minlength=n_rows)
    del out_row
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=col.device)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr.int(), col


def build_party_csr(torch, n_local, n_edges, n_parties, rank, seed, device, own_first=False):
    """CSR-by-destination of the edges party `rank` owns.  Sources are its local vertices (row index 0..n_local),
    destinations are global vertices grouped by owner: output row of global vertex v = (v % P) * n_local + v // P.
    own_first: the owner blocks are rotated so that the party's own block comes first and the block of party rank + j is the
    j-th (the order in which the fused step completes and signals them)."""
    n_global = n_local * n_parties
    src, dst = rmat_edges(torch, n_global, n_edges, seed + 1000 * rank, device)
    src = src % n_local  # the party's own vertices (local row index)
    owner = dst % n_parties
    if own_first:
        owner = (owner - rank) % n_parties
    out_row = owner * n_local + dst // n_parties
    del dst, owner
    return csr_from_edges(torch, src, out_row, n_local, n_global)


def build_uniform_csr(torch, n_local, n_edges, seed, device):
    """The no-hub control: sources and destinations uniform over the vertices, so a row read is almost never served by L2 and
    the algorithmic bytes of SURVEY 8d are (nearly) DRAM bytes."""
    g = torch.Generator(device=device).manual_seed(seed)
    src = torch.randint(0, n_local, (n_edges,), dtype=torch.int64, device=device, generator=g)
    dst = torch.randint(0, n_local, (n_edges,), dtype=torch.int64, device=device, generator=g)
    return csr_from_edges(torch, src, dst, n_local, n_local)


def exchange_and_sum(dist, y, recv, v, n_parties, n_local, D, add):
    """Mirror-update exchange of one GAS iteration (ssk.h:835 -> 1067/1090) in its plainest form: block i of the dense y
    goes to party i, then the received blocks are summed share-locally (GatherComp add, gcn.h:456).  `add(a, b, out)` is the
    engine's add.  The pull exchange below must reproduce exactly this."""
    dist.all_to_all_single(recv, y)
    blocks = recv.view(n_parties, n_local, D)
    add(blocks[0], blocks[1], v)
    for j in range(2, n_parties):
        add(v, blocks[j], v)
    return v


class RawCuda:
    """A cgb_malloc allocation viewed as a torch tensor (cudaMalloc memory is IPC-exportable, torch's is not)."""

    def __init__(self, ptr, shape, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


# ------------------------------------------------------------------------------------------------------------
# N > 1: the pull exchange
# ------------------------------------------------------------------------------------------------------------
FLAG_BYTES = 4096  # ready[P * S] | ack[P * S] | err at the start of every rank's exported allocation (P * S <= 256)


def pull_orders(rank, P):
    """Producer order (destination of the j-th remote gather) and consumer order (source whose block is pulled j-th): the
    block for party rank+j is gathered j-th, so the block FROM party rank-j is that party's j-th too and arrives j-th."""
    return [(rank + j) % P for j in range(1, P)], [(rank - j) % P for j in range(1, P)]


def sub_blocks(P):
    """Row ranges each remote block is cut into (as fractions of the destination party's rows).  A pull can only start when its
    block is complete, so with few parties the blocks are cut finer: the pull of one piece overlaps the gather of the next and
    only the last piece's pull is exposed (2 parties: 4 pieces of the one remote block; 8 parties: the 7 blocks as they are)."""
    total = int(os.environ.get("CGB_PIECES_TOTAL", "8"))
    return max(1, min(total // P, 31 // max(1, P - 1)))


def exchange_index_lists(torch, dist, my_lists, rank, P):
    """my_lists[t] = rows of party t that have an edge from me (ascending; my PosVec for t, ssk.h:507-516).  Returns
    got[s] = rows of MINE that have an edge from party s.  Variable sizes: counts first, then point-to-point transfers (works on
    NCCL and on gloo, which the CPU test uses)."""
    sizes = [None] * P
    dist.all_gather_object(sizes, [int(l.numel()) for l in my_lists])
    out = [torch.empty(sizes[s][rank], dtype=torch.int32, device=my_lists[0].device) for s in range(P)]
    out[rank] = my_lists[rank].clone()
    ops = []
    for j in range(1, P):
        t, s = (rank + j) % P, (rank - j) % P
        if sizes[rank][t]:
            ops.append(dist.P2POp(dist.isend, my_lists[t].contiguous(), t))
        if sizes[s][rank]:
            ops.append(dist.P2POp(dist.irecv, out[s], s))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    return out, sizes


def setup_pull(torch, dist, ctx, rowptr, col, rank, P, n_local, D, dev):
    import cognn_b200

    S = sub_blocks(P)
    cuts = [n_local * u // S for u in range(S + 1)]

    def sub_csr(lo, hi):
        e0, e1 = int(rowptr[lo]), int(rowptr[hi])
        return ctx.csr_create((rowptr[lo:hi + 1] - rowptr[lo]).int().contiguous(), col[e0:e1].contiguous(), n_local)

    own_csr = sub_csr(rank * n_local, (rank + 1) * n_local)
    csrs = {(t, u): sub_csr(t * n_local + cuts[u], t * n_local + cuts[u + 1]) for t in range(P) if t != rank for u in range(S)}
    # PosVec of preprocessing: per destination party the rows (of THAT party) that receive an edge from me, piece by piece
    empty = torch.empty(0, dtype=torch.int32, device=dev)
    nz_mine = [empty if t == rank else torch.cat([csrs[(t, u)].nonempty_rows() + cuts[u] for u in range(S)]) for t in range(P)]
    counts_mine = [[0] * S if t == rank else [csrs[(t, u)].n_nonempty for u in range(S)] for t in range(P)]
    nz_from, _ = exchange_index_lists(torch, dist, nz_mine, rank, P)
    all_counts = [None] * P
    dist.all_gather_object(all_counts, counts_mine)  # all_counts[s][t][u]: rows of piece u of the block s -> t
    # one exported allocation per rank: flags, then two staging buffers (step parity) per remote piece
    offs, cur = {}, FLAG_BYTES
    for b in range(2):
        for (t, u), c in csrs.items():
            offs[(b, t, u)] = cur
            cur += (c.n_nonempty * D * 8 + 255) & ~255
    base = ctx.malloc(max(cur, FLAG_BYTES))
    ctx.check(ctx.lib.cgb_memset(ctx.handle, base, 0, FLAG_BYTES))
    ctx.sync()
    everyone = [None] * P
    dist.all_gather_object(everyone, (ctx.ipc_export(base), offs))
    peer_base = [base if s == rank else ctx.ipc_open(everyone[s][0]) for s in range(P)]
    side = torch.cuda.Stream(device=dev, priority=-1)  # pull CTAs are placed ahead of the queued gather CTAs
    with torch.cuda.stream(side):
        ctx_side = cognn_b200.Context(dev.index)
    # per source s and piece u: the slice of nz_from[s] that belongs to the piece
    nz_piece = {}
    for s in range(P):
        if s == rank:
            continue
        o = 0
        for u in range(S):
            n = all_counts[s][rank][u]
            nz_piece[(s, u)] = (nz_from[s][o:o + n].contiguous(), n)
            o += n
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    return {"S": S, "own_csr": own_csr, "csrs": csrs, "nz_piece": nz_piece, "base": base, "offs": offs, "peer_base": peer_base,
            "peer_offs": [everyone[s][1] for s in range(P)], "side": side, "ctx_side": ctx_side, "step": 0,
            "wait_mode": 1 if os.environ.get("CGB_FLAG_WAIT", "memop") == "spin" else 0,
            "pull_ctas": int(os.environ.get("CGB_PULL_CTAS", "128")),
            "t0": ev(), "own": ev(), "gat": {k: ev() for k in csrs}, "add": {k: ev() for k in nz_piece}, "end": ev(),
            "err": torch.as_tensor(RawCuda(base + 8 * P * S, (1,), "<i4"), device=dev),
            "bytes_out": sum(c.n_nonempty for c in csrs.values()) * D * 8,
            "rows_out": {k: c.n_nonempty for k, c in csrs.items()}}


def pull_step(torch, ctx, x, v, pl, rank, P, D):
    """One step: own block gathered densely into v; every remote block gathered piece by piece in compact form into staging
    buffers its consumer has mapped, a flag raised in the consumer's memory after each piece; meanwhile the side stream waits
    for the other parties' flags in arrival order and pulls + adds their pieces (cgb_scatter_add_rows over NVLink).  No
    collective, no dense block copy, nothing lands in local memory before it is added."""
    pl["step"] += 1
    i = pl["step"]
    b, S = i & 1, pl["S"]
    main, side, cs = torch.cuda.current_stream(), pl["side"], pl["ctx_side"]
    prod, cons = pull_orders(rank, P)
    ready = lambda r, src, u: pl["peer_base"][r] + 4 * (src * S + u)            # noqa: E731  ready[src, u] in rank r's memory
    ack = lambda r, dst, u: pl["peer_base"][r] + 4 * (P * S + dst * S + u)      # noqa: E731  ack[dst, u] in rank r's memory
    err = pl["base"] + 8 * P * S
    pl["t0"].record(main)
    ctx.gather_sum(pl["own_csr"], x, None, out=v)  # every row of v is written (zeros where no local edge ends)
    pl["own"].record(main)
    side.wait_event(pl["own"])
    for t in prod:
        for u in range(S):
            if i > 2:  # the buffer of this parity was read by t two steps ago: its ack must have arrived
                ctx.flag_wait(ack(rank, t, u), i - 2, pl["wait_mode"], err)
            ctx.gather_sum_compact(pl["csrs"][(t, u)], x, out_ptr=pl["base"] + pl["offs"][(b, t, u)])
            ctx.flag_signal(ready(t, rank, u), i)
            pl["gat"][(t, u)].record(main)
    with torch.cuda.stream(side):
        for s in cons:
            for u in range(S):
                idx, n = pl["nz_piece"][(s, u)]
                cs.flag_wait(ready(rank, s, u), i, pl["wait_mode"], err)
                cs.scatter_add_rows(idx, pl["peer_base"][s] + pl["peer_offs"][s][(b, rank, u)], v, n=n, D=D, n_ctas=pl["pull_ctas"])
                cs.flag_signal(ack(s, rank, u), i)
                pl["add"][(s, u)].record(side)
        pl["end"].record(side)
    main.wait_event(pl["end"])
    return v


def pull_phases(pl, rank, P):
    """Timeline of the LAST step on this rank (ms after the step's start), read after a synchronize."""
    t0, S = pl["t0"], pl["S"]
    prod, cons = pull_orders(rank, P)
    return {"own_gather_done": round(t0.elapsed_time(pl["own"]), 3),
            "remote_gather_done": [round(t0.elapsed_time(pl["gat"][(t, u)]), 3) for t in prod for u in range(S)],
            "piece_added": [round(t0.elapsed_time(pl["add"][(s, u)]), 3) for s in cons for u in range(S)],
            "step_done": round(t0.elapsed_time(pl["end"]), 3), "pieces_per_block": S,
            "rows_sent_per_piece": [pl["rows_out"][(t, u)] for t in prod for u in range(S)], "nvlink_bytes_out": pl["bytes_out"]}


def setup_fused(torch, dist, ctx, rank, P, n_local, n_edges, D, dev, v):
    """The fused step (cgb_gather_sum_signal): ONE CSR over all of the party's out-edges with the destination blocks in the order
    [own | rank+1 | rank+2 | ...], remote blocks cut into sub_blocks(P) row pieces; every piece has a staging buffer (two step
    parities) its consumer has mapped and a flag in the consumer's memory that the gather kernel itself raises."""
    import cognn_b200

    S = sub_blocks(P)
    cuts = [n_local * u // S for u in range(S + 1)]
    rowptr, col = build_party_csr(torch, n_local, n_edges, P, rank, 42, dev, own_first=True)
    csr = ctx.csr_create(rowptr, col, n_local)
    del rowptr, col
    nz = csr.nonempty_rows().long()
    prod, cons = pull_orders(rank, P)
    offsets, pieces = [0], []          # pieces[i] = (dest party t, piece u) of block i + 1
    for j, t in enumerate(prod, start=1):
        for u in range(S):
            offsets.append(j * n_local + cuts[u])
            pieces.append((t, u))
    offsets.append(P * n_local)        # block 0 = own rows [0, n_local), then the remote pieces in signalling order
    bounds = torch.searchsorted(nz, torch.tensor(offsets, device=dev))
    bl = bounds.tolist()
    counts_mine = [[0] * S for _ in range(P)]
    nz_mine = [torch.empty(0, dtype=torch.int32, device=dev) for _ in range(P)]
    for i, (t, u) in enumerate(pieces, start=1):
        counts_mine[t][u] = bl[i + 1] - bl[i]
    for j, t in enumerate(prod, start=1):  # rows of party t (its local index) that receive an edge from me: my PosVec for t
        lo, hi = bl[1 + (j - 1) * S], bl[1 + j * S]
        nz_mine[t] = (nz[lo:hi] - j * n_local).int()
    nz_from, _ = exchange_index_lists(torch, dist, nz_mine, rank, P)
    all_counts = [None] * P
    dist.all_gather_object(all_counts, counts_mine)
    offs, cur = {}, FLAG_BYTES
    for b in range(2):
        for (t, u) in pieces:
            offs[(b, t, u)] = cur
            cur += (counts_mine[t][u] * D * 8 + 255) & ~255
    base = ctx.malloc(max(cur, FLAG_BYTES))
    ctx.check(ctx.lib.cgb_memset(ctx.handle, base, 0, FLAG_BYTES))
    ctx.sync()
    everyone = [None] * P
    dist.all_gather_object(everyone, (ctx.ipc_export(base), offs))
    peer_base = [base if s == rank else ctx.ipc_open(everyone[s][0]) for s in range(P)]
    side = torch.cuda.Stream(device=dev, priority=-1)  # pull CTAs are placed ahead of the queued gather CTAs
    with torch.cuda.stream(side):
        ctx_side = cognn_b200.Context(dev.index)
    nz_piece = {}
    for s in cons:
        o = 0
        for u in range(S):
            n = all_counts[s][rank][u]
            nz_piece[(s, u)] = (nz_from[s][o:o + n].contiguous(), n)
            o += n
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    return {"S": S, "csr": csr, "offsets": offsets, "pieces": pieces, "nz_piece": nz_piece, "base": base, "offs": offs,
            "peer_base": peer_base, "peer_offs": [everyone[s][1] for s in range(P)], "side": side, "ctx_side": ctx_side, "step": 0,
            "wait_mode": 1 if os.environ.get("CGB_FLAG_WAIT", "memop") == "spin" else 0,
            "pull_ctas": int(os.environ.get("CGB_PULL_CTAS", "128")),
            "t0": ev(), "launched": ev(), "add": {k: ev() for k in nz_piece}, "own_seen": ev(), "end": ev(),
            "err": torch.as_tensor(RawCuda(base + 8 * P * S, (1,), "<i4"), device=dev),
            "bytes_out": sum(counts_mine[t][u] for (t, u) in pieces) * D * 8,
            "rows_out": {k: counts_mine[k[0]][k[1]] for k in pieces}}


def fused_step(torch, ctx, x, v, fz, rank, P, D):
    """One step = ONE gather launch.  The kernel stores the own block densely into v and every remote piece in compact form into
    its staging buffer and raises each block's flag the moment that block is complete; the side stream (high priority) waits for
    the own block's flag, then for the other parties' flags in arrival order, and pulls + adds their pieces over NVLink."""
    fz["step"] += 1
    i = fz["step"]
    b, S = i & 1, fz["S"]
    main, side, cs = torch.cuda.current_stream(), fz["side"], fz["ctx_side"]
    _, cons = pull_orders(rank, P)
    ready = lambda r, src, u: fz["peer_base"][r] + 4 * (src * S + u)            # noqa: E731  ready[src, u] in rank r's memory
    ack = lambda r, dst, u: fz["peer_base"][r] + 4 * (P * S + dst * S + u)      # noqa: E731  ack[dst, u] in rank r's memory
    err = fz["base"] + 8 * P * S
    fz["t0"].record(main)
    if i > 2:  # the staging buffers of this parity were read two steps ago: those acks must have arrived
        for (t, u) in fz["pieces"]:
            ctx.flag_wait(ack(rank, t, u), i - 2, fz["wait_mode"], err)
    bases = [v.data_ptr()] + [fz["base"] + fz["offs"][(b, t, u)] for (t, u) in fz["pieces"]]
    flags = [ready(rank, rank, 0)] + [ready(t, rank, u) for (t, u) in fz["pieces"]]
    ctx.gather_sum_signal(fz["csr"], x, fz["offsets"], bases, [False] + [True] * len(fz["pieces"]), flags, i)
    fz["launched"].record(main)
    with torch.cuda.stream(side):
        cs.flag_wait(ready(rank, rank, 0), i, fz["wait_mode"], err)  # own rows are in v: the additions may start
        fz["own_seen"].record(side)
        for s in cons:
            for u in range(S):
                idx, n = fz["nz_piece"][(s, u)]
                cs.flag_wait(ready(rank, s, u), i, fz["wait_mode"], err)
                cs.scatter_add_rows(idx, fz["peer_base"][s] + fz["peer_offs"][s][(b, rank, u)], v, n=n, D=D, n_ctas=fz["pull_ctas"])
                cs.flag_signal(ack(s, rank, u), i)
                fz["add"][(s, u)].record(side)
        fz["end"].record(side)
    main.wait_event(fz["end"])
    return v


def fused_phases(fz, rank, P):
    t0, S = fz["t0"], fz["S"]
    _, cons = pull_orders(rank, P)
    return {"gather_kernel_done": round(t0.elapsed_time(fz["launched"]), 3), "own_block_flag_seen": round(t0.elapsed_time(fz["own_seen"]), 3),
            "piece_added": [round(t0.elapsed_time(fz["add"][(s, u)]), 3) for s in cons for u in range(S)],
            "step_done": round(t0.elapsed_time(fz["end"]), 3), "pieces_per_block": S,
            "rows_sent_per_piece": [fz["rows_out"][k] for k in fz["pieces"]], "nvlink_bytes_out": fz["bytes_out"]}


# ------------------------------------------------------------------------------------------------------------
# secure-GCN epoch records (the other half of BASELINE.json's metric)
# ------------------------------------------------------------------------------------------------------------
def secure_gcn_epoch_probe(timeout_s=300):
    """N = 1: BASELINE.json configs[0], the 2-party CoGNN-Opt training epoch on the synthetic Cora-shaped graph, both parties on
    this GPU (loopback plane).  Child process, so nothing it does can disturb the gather measurement; any failure is reported,
    never raised.  `--check` compares the final weight shares with the epoch oracle."""
    cmd = [sys.executable, os.path.join(ROOT, "tools", "epoch_bench.py"), "--shape", "cora", "--parties", "2", "--epochs", "4", "--cpu",
           "--check"]
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        rec = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        keep = ("shape", "parties", "plane", "N", "E", "cfg", "iterations", "online_s", "online_gpu_s_this_rank", "online_mode", "offline_dealer_s", "launches",
                "rounds", "graph_replays", "cpu_oracle", "bit_exact_vs_oracle", "note")
        out = {k: rec[k] for k in keep if k in rec}
        out["unit"] = "s per epoch (online phase; offline = trusted-dealer emulation, reported beside it)"
        return out
    except Exception as ex:  # noqa: BLE001
        return {"error": f"{type(ex).__name__}: {ex}"[:300]}


EPOCH_SHAPE = {2: ("cora", None, "configs[0]"), 4: ("citeseer", 0.1, "configs[2] (optimize-gcn arm)"), 8: ("arxiv", None, "configs[3]")}


def secure_gcn_epoch_nccl(rank, P, timeout_s=420):
    """N = 2 / 4 / 8: the training epoch of the configuration BASELINE.json names for that party count, one party per GPU over
    the engine's own NCCL plane, every hosted share checked against the epoch oracle on every rank (`--check`).  Every rank
    starts tools/epoch_bench.py as a CHILD with the same rank / world and the next rendezvous port: a failure or a hang in there
    costs the record, never the bench line."""
    shape, inter, which = EPOCH_SHAPE[P]
    cmd = [sys.executable, os.path.join(ROOT, "tools", "epoch_bench.py"), "--shape", shape, "--parties", str(P), "--epochs", "3", "--check"]
    if inter is not None:
        cmd += ["--inter", str(inter)]
    env = dict(os.environ)
    env["MASTER_PORT"] = str(int(os.environ.get("MASTER_PORT", "29500")) + 1)
    env["MASTER_ADDR"] = os.environ.get("MASTER_ADDR", "127.0.0.1")
    for k in ("TORCHELASTIC_RUN_ID", "TORCHELASTIC_USE_AGENT_STORE", "TORCHELASTIC_RESTART_COUNT", "TORCHELASTIC_MAX_RESTARTS"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        if rank != 0:
            return None
        rec = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        keep = ("shape", "parties", "plane", "N", "E", "inter_party_edges", "cfg", "iterations", "online_s", "online_gpu_s_this_rank", "online_mode",
                "offline_dealer_s", "launches", "rounds", "graph_replays", "bit_exact_vs_oracle", "checked", "note")
        out = {"config": which}
        out.update({k: rec[k] for k in keep if k in rec})
        out["unit"] = "s per epoch (online phase, max over ranks; offline = trusted-dealer emulation, reported beside it)"
        return out
    except Exception as ex:  # noqa: BLE001
        return {"config": which, "error": f"{type(ex).__name__}: {ex}"[:300]} if rank == 0 else None


# ------------------------------------------------------------------------------------------------------------
# Beaver matmul record (configs[4], second half): 2^20 x F . F x H on both pipes
# ------------------------------------------------------------------------------------------------------------
def matmul_record(torch, ctx, dev, shapes=((128, 128), (256, 256), (512, 512), (128, 512), (512, 128)), M=1 << 20, reps=3):
    import numpy as np

    imad_peak, tensor_peak = ctx.probe_imad_peak(), ctx.probe_tensor_i8_peak()
    g = torch.Generator(device=dev).manual_seed(7)
    rows = []
    for F, H in shapes:
        A = torch.randint(-2**63, 2**63 - 1, (M, F), dtype=torch.int64, device=dev, generator=g)
        B = torch.randint(-2**63, 2**63 - 1, (F, H), dtype=torch.int64, device=dev, generator=g)
        C = torch.empty((M, H), dtype=torch.int64, device=dev)
        pick = torch.randint(0, M, (48,), device=dev, generator=g)
        want = A[pick].cpu().numpy().view(np.uint64) @ B.cpu().numpy().view(np.uint64)  # numpy integer matmul wraps mod 2^64
        row = {"F": F, "H": H, "u64_mac": M * F * H}
        for impl in ("imad", "tc"):
            ctx.set_matmul_impl(impl)
            ctx.matmul(A, B, out=C)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = None
            for _ in range(reps):
                e0.record()
                ctx.matmul(A, B, out=C)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            rate = M * F * H / (best * 1e-3)
            ok = bool(np.array_equal(C[pick].cpu().numpy().view(np.uint64), want))
            row[impl] = {"ms": best, "u64_mac_per_s": rate, "bit_exact_sampled_rows": ok, "kernel": ctx.last_kernel,
                         "frac_of_pipe": rate / imad_peak if impl == "imad" else rate * 36.0 / tensor_peak}
        ctx.set_matmul_impl("auto")
        ctx.matmul(A, B, out=C)
        row["auto_picks"] = "tc" if "matmul_tc" in ctx.last_kernel else "imad"
        rows.append(row)
        del A, B, C
    ctx.set_matmul_impl(None)
    return {"workload": f"configs[4]: Beaver matmul local term, {M} x F . F x H, u64 mod 2^64, operands uniform random",
            "pipe_peaks_measured_here": {"imad_u64_mac_per_s": imad_peak, "tensor_u8_limb_mac_per_s": tensor_peak,
                                         "how": "cgb_probe_imad_peak (register operands) / cgb_probe_tensor_i8_peak (the limb kernel's "
                                                "tcgen05.mma mix on resident shared-memory operands)"},
            "frac_of_pipe": "imad: u64 MAC/s over the IMAD ceiling; tc: 36 limb MACs per u64 MAC over the tensor ceiling; times include "
                            "the limb-split pre-pass",
            "shapes": rows}


def algorithmic_bytes(n_rows, n_edges, D):
    # SURVEY.md 8d, fused SpMM form: every edge = one 8*D-byte row read + a 4-byte index, no cache-reuse credit
    return (8 * D + 4) * n_edges + 4 * (n_rows + 1) + 8 * D * n_rows


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for (_, l) in self.lines]
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for nme, val in zip(names, p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(E, D, P):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of THIS configuration
    (profiles/ncu_traffic.json), or None: only the N = 1 default workload has one, other lines say null."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if P != 1 or not os.path.exists(p):
        return None, None
    try:
        rec = json.load(open(p)).get("gather_sum_kernel", {})
        if int(rec.get("edges", 100_000_000)) != E:
            return None, None
        return rec.get(f"D{D}_bytes_per_launch"), rec.get("source")
    except Exception:
        return None, None


def host_threads():
    """All host cores for the CPU legs: torchrun exports OMP_NUM_THREADS=1 to every rank, which must not cap the baseline."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_full_pass(pyoracle, rp, cl, xh, min_seconds, reps):
    best, total, n, y = None, 0.0, 0, None
    while n < reps or (total < min_seconds and n < 200):
        t0 = time.perf_counter()
        y = pyoracle.gather_sum_csr(rp, cl, xh)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        total += dt
        n += 1
    return best, total, n, y


def cpu_sample(torch, rowptr, col, x, D, frac_rows, reps=3, threads=None, min_seconds=10.0):
    """Times the CPU oracle on the first `frac_rows` destination rows of the SAME graph (same index skew, same random row
    reads into the full x).  Returns (edges/s, description, threads, best seconds, rows, result)."""
    import numpy as np

    from oracle import pyoracle

    pyoracle.build()
    pyoracle.set_num_threads(threads or host_threads())
    n_rows = rowptr.numel() - 1
    rows = max(1, int(n_rows * frac_rows))
    rp = rowptr[: rows + 1].cpu().numpy().view(np.uint32).copy()
    e = int(rp[-1])
    cl = col[:e].cpu().numpy().view(np.uint32).copy()
    xh = x.cpu().numpy().view(np.uint64)
    best, total, n, y = cpu_full_pass(pyoracle, rp, cl, xh, min_seconds, reps)
    return e / best, (f"first {rows} of {n_rows} destination rows ({e} edges) of the same graph, best of {n} passes "
                      f"({total:.1f} s of CPU work)"), pyoracle.num_threads(), best, rows, y


def reference_arm(args, torch, rank, world, config, K, D, E, n_local):
    """The reference's CPU path (oracle port) on the same workload: P parties' gathers and the block sums, all host cores."""
    import numpy as np

    from oracle import pyoracle

    if rank != 0:
        return 0
    P = world if world > 1 else 1
    pyoracle.build()
    pyoracle.set_num_threads(host_threads())
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    if dev == "cpu":  # local smoke only
        E_ref = min(E, 2_000_000)
        n_ref = max(1, E_ref // 16)
    else:
        E_ref, n_ref = E, n_local
    rows = max(1, int(n_ref * P * args.cpu_frac))
    parts = []
    for p in range(P):
        rowptr, col = build_party_csr(torch, n_ref, E_ref, P, p, 42, dev)
        g = torch.Generator(device=dev).manual_seed(43 + p)
        x = torch.randint(-2**63, 2**63 - 1, (n_ref, D), dtype=torch.int64, device=dev, generator=g)
        rp = rowptr[: rows + 1].cpu().numpy().view(np.uint32).copy()
        e = int(rp[-1])
        parts.append((rp, col[:e].cpu().numpy().view(np.uint32).copy(), x.cpu().numpy().view(np.uint64), e))
        del rowptr, col, x
    edges_per_step = sum(p[3] for p in parts)
    V = np.zeros((rows, D), dtype=np.uint64)  # row (t, i) = vertex i of party t: all parties' received sums side by side

    def step():
        # party 0's gather initialises the sums, every further party's gather accumulates its blocks onto them: the P gathers
        # and the (P - 1) block additions per destination party of the GPU arm, fused the way a CPU would do it
        pyoracle.gather_sum_csr(parts[0][0], parts[0][1], parts[0][2], out=V)
        for rp, cl, xh, _ in parts[1:]:
            pyoracle.gather_sum_csr(rp, cl, xh, delta=V, out=V)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    dt = time.perf_counter() - t0
    val = edges_per_step * K / dt
    sample = (f"each step = all {P} part{'y' if P == 1 else 'ies'}: first {rows} of {n_ref * P} destination rows of each party's graph "
              f"({edges_per_step} edges per step), gathers accumulated into the per-party sums in place")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": args.warmup, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": pyoracle.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU oracle port of the reference path (reference unbuildable here, SURVEY.md 8c); host cores only; the same "
                    "P-party work as the CUDA arm (P gathers of E edges + block sums), aggregate edges/s"}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--edges", type=int, default=100_000_000, help="edges per party (per GPU)")
    ap.add_argument("--dim", type=int, default=16, help="u64 columns per share row (hidden_dim)")
    ap.add_argument("--cpu-frac", type=float, default=1.0, help="fraction of rows in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-epoch", action="store_true", help="skip the secure-GCN epoch record")
    ap.add_argument("--no-matmul", action="store_true", help="skip the Beaver matmul record (N = 1 only)")
    ap.add_argument("--no-uniform", action="store_true", help="skip the uniform-graph control of the roofline (N = 1 only)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "pull", "nccl"],
                    help="N > 1: fused = ONE gather launch that stores compact blocks and raises per-block flags in the consumers' "
                         "memory, consumers pull + add over NVLink (default); pull = the same with one gather launch per block and "
                         "stream-ordered flags; nccl = dense gather then all_to_all + sums")
    args = ap.parse_args()

    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    K = max(args.steps, 1)
    D = args.dim
    E = args.edges
    n_local = max(1, E // 16)
    P = world if world > 1 else 1
    config = {
        "workload": f"configs[4] kernel sweep: fused share gather-sum, RMAT(0.57,0.19,0.19) power-law graph, "
                    f"{E} edges and {n_local} vertices per party, D={D} u64 columns, {P} part{'y' if P == 1 else 'ies'}",
        "edges_per_party": E, "vertices_per_party": n_local, "D": D, "parties": P,
        "l2": "inputs larger than L2 (share rows + indices >> 126 MB); no flush needed",
        "seed": 42,
        "exchange": "none (single party)" if P == 1 else
        "mirror-update blocks of every ordered pair of parties, summed into the destination party's vertex rows (ssk.h:835 -> 1090, "
        "gcn.h:456)",
    }

    if args.impl == "reference":
        return reference_arm(args, torch, rank, world, config, K, D, E, n_local)

    # -------------------------------------------------------------------------------------------------------
    import cognn_b200

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    ctx = cognn_b200.Context(local_rank)

    rowptr, col = build_party_csr(torch, n_local, E, P, rank, 42, dev)
    g = torch.Generator(device=dev).manual_seed(43 + rank)
    x = torch.randint(-2**63, 2**63 - 1, (n_local, D), dtype=torch.int64, device=dev, generator=g)
    add = lambda a, b, o: ctx.add(a, b, out=o)  # noqa: E731
    exchange_impl, exchange_check, pl, fz = None, None, None, None
    if P == 1:
        csr = ctx.csr_create(rowptr, col, n_local)
        y = torch.empty((n_local, D), dtype=torch.int64, device=dev)
        v = None
    else:
        v = torch.empty((n_local, D), dtype=torch.int64, device=dev)
        # reference result of the step, outside the timed region: dense gather, NCCL all-to-all, block sums
        csr = ctx.csr_create(rowptr, col, n_local)
        y = torch.empty((n_local * P, D), dtype=torch.int64, device=dev)
        recv = torch.empty_like(y)
        ctx.gather_sum(csr, x, None, out=y)
        want_v = exchange_and_sum(dist, y, recv, torch.empty_like(v), P, n_local, D, add).clone()
        exchange_impl = args.exchange
        if exchange_impl in ("pull", "fused"):
            del recv, y
            csr.destroy()
            csr = y = recv = None
            torch.cuda.empty_cache()
            try:  # peer windows need CUDA IPC between the ranks' processes
                if exchange_impl == "pull":
                    pl = setup_pull(torch, dist, ctx, rowptr, col, rank, P, n_local, D, dev)
                else:
                    del rowptr, col
                    torch.cuda.empty_cache()
                    fz = setup_fused(torch, dist, ctx, rank, P, n_local, E, D, dev, v)
                ok = torch.ones(1, dtype=torch.int32, device=dev)
            except Exception as ex:  # noqa: BLE001
                sys.stderr.write(f"[bench] rank {rank}: peer memory unavailable ({ex}); falling back to --exchange nccl\n")
                ok = torch.zeros(1, dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                exchange_impl, pl, fz = "nccl", None, None
                rowptr, col = build_party_csr(torch, n_local, E, P, rank, 42, dev)
                csr = ctx.csr_create(rowptr, col, n_local)
                y = torch.empty((n_local * P, D), dtype=torch.int64, device=dev)
                recv = torch.empty_like(y)
        if exchange_impl in ("pull", "fused"):
            good = 1
            for _ in range(3):  # both staging parities and the ack path
                if exchange_impl == "pull":
                    pull_step(torch, ctx, x, v, pl, rank, P, D)
                else:
                    fused_step(torch, ctx, x, v, fz, rank, P, D)
                torch.cuda.synchronize()
                good &= int(torch.equal(v, want_v))
            good &= int(int((pl or fz)["err"].item()) == 0)
            t = torch.tensor([good], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            exchange_check = {"against": "dense gather + NCCL all_to_all_single + block sums of the same inputs", "steps": 3,
                              "ok": bool(int(t.item()))}
            assert exchange_check["ok"], "fused / pull exchange differs from the gather + all-to-all + sum path"
            dist.barrier()
        config["exchange_impl"] = {
            "fused": "ONE gather launch per step over all out-edges (cgb_gather_sum_signal): own block stored densely, remote blocks "
                     "(cut into max(1, 8 // parties) row pieces) in compact form (rows of destinations that have an edge, the PosVec "
                     "both sides hold) into IPC-mapped staging; the kernel itself raises each piece's arrival flag in the consumer's "
                     "memory when the piece is complete; the consumer's side stream waits per piece and pulls + adds it over NVLink "
                     "(cgb_scatter_add_rows) while the producer is still gathering; no collective in the step",
            "pull": "one gather per destination party (remote blocks cut into max(1, 8 // parties) row pieces); remote pieces in compact "
                    "form (rows of destinations that have an edge, the PosVec both sides hold) into IPC-mapped staging, arrival flag "
                    "raised in the consumer's memory; the consumer's side stream waits per piece and pulls + adds it over NVLink "
                    "(cgb_scatter_add_rows); no collective in the step",
            "nccl": "dense gather, NCCL all_to_all_single of the N_p x D blocks, cgb_add sums"}[exchange_impl]

    step_no = [0]

    def step():
        step_no[0] += 1
        if P == 1:
            ctx.gather_sum(csr, x, None, out=y)
        elif exchange_impl == "pull":
            pull_step(torch, ctx, x, v, pl, rank, P, D)
        elif exchange_impl == "fused":
            fused_step(torch, ctx, x, v, fz, rank, P, D)
        else:
            ctx.gather_sum(csr, x, None, out=y)
            exchange_and_sum(dist, y, recv, v, P, n_local, D, add)

    def barrier():
        if P > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    all_ctx = [ctx] + ([(pl or fz)["ctx_side"]] if (pl or fz) else [])
    launches0 = sum(c.launches for c in all_ctx)
    t_wall0 = time.time()
    torch.cuda.profiler.start()  # ncu --profile-from-start off captures exactly the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        kev[i][0].record()
        step()
        kev[i][1].record()
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    t_wall1 = time.time()
    launches = sum(c.launches for c in all_ctx) - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / K
    kernel_name = ctx.last_kernel
    phases = None
    if fz is not None:
        if int(fz["err"].item()) != 0:
            raise RuntimeError("a flag wait timed out inside the timed region")
        allp = [None] * P
        dist.all_gather_object(allp, fused_phases(fz, rank, P))
        phases = {"per_rank_last_step_timeline_ms": allp}
        # the gather launch on its own (no flags, no pulls running beside it), for the roofline of the dominant kernel
        scr = [torch.empty((max(1, fz["rows_out"][k]), D), dtype=torch.int64, device=dev) for k in fz["pieces"]]
        bases = [v.data_ptr()] + [t.data_ptr() for t in scr]
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        ctx.gather_sum_signal(fz["csr"], x, fz["offsets"], bases, [False] + [True] * len(scr), [0] * len(bases), 0)
        ka.record()
        for _ in range(reps):
            ctx.gather_sum_signal(fz["csr"], x, fz["offsets"], bases, [False] + [True] * len(scr), [0] * len(bases), 0)
        kb.record()
        torch.cuda.synchronize()
        kernel_ms = ka.elapsed_time(kb) / reps
        kernel_name = "gather_chunk_signal_kernel<VEC=4,LANES=4,U=4,128,1024,IPL=2> (256-bit row loads, per-block completion flags)"
        del scr
    if pl is not None:
        if int(pl["err"].item()) != 0:
            raise RuntimeError("a flag wait timed out inside the timed region")
        allp = [None] * P
        dist.all_gather_object(allp, pull_phases(pl, rank, P))
        phases = {"per_rank_last_step_timeline_ms": allp}
        # the P gather launches of one step on their own (no waits, no pulls), for the roofline of the dominant kernel
        scratch = torch.empty((max(pl["rows_out"].values()), D), dtype=torch.int64, device=dev)
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        ka.record()
        for _ in range(reps):
            ctx.gather_sum(pl["own_csr"], x, None, out=v)
            for k, c in pl["csrs"].items():
                ctx.gather_sum_compact(c, x, out=scratch[: pl["rows_out"][k]])
        kb.record()
        torch.cuda.synchronize()
        kernel_ms = ka.elapsed_time(kb) / reps
        del scratch
    if P > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = E * P / (ms_per_step * 1e-3)

    peak, peak_src = measured_peak_hbm()
    n_out_rows = n_local if P == 1 else (n_local + sum((pl or fz)["rows_out"].values()) if (pl or fz) else n_local * P)
    n_launch = 1 + len(pl["csrs"]) if pl else 1
    alg = algorithmic_bytes(n_out_rows, E, D)
    achieved = alg / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(E, D, P)
    roofline = {"bound": "hbm", "kernel": kernel_name + (f" x {n_launch} launches per step" if n_launch > 1 else ""),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "kernel_ms": kernel_ms,
                "frac_of_8TBs_nominal": achieved / 8000.0,
                "dram_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "traffic_source": traffic_src,
                "note": "`achieved` uses SURVEY 8d's algorithmic bytes, which charge every edge a full 8*D-byte row read (no cache-reuse "
                        "credit); hub rows are served by L2, so `frac` can exceed 1 and is NOT an HBM fraction.  `dram_frac` = measured "
                        "DRAM bytes (ncu `traffic`) / kernel_ms / peak is; `uniform_graph` repeats the kernel on a hub-free graph where "
                        "the algorithmic bytes are (nearly) DRAM bytes.  The kernel's binding limit on this power-law workload is the L2 "
                        "slice throughput (DESIGN.md section 5)"}

    # ---- parity of the timed configuration + CPU baseline (rank 0, N == 1) -------------------------------------------------
    parity, cpu_baseline = None, None
    if rank == 0 and P == 1 and not args.no_cpu_baseline:
        import numpy as np

        val, sample, cores, secs, rows, y_cpu = cpu_sample(torch, rowptr, col, x, D, args.cpu_frac)
        y_gpu = y[:rows].cpu().numpy().view(np.uint64)
        ok = bool(np.array_equal(y_gpu, y_cpu))
        parity = {"checked_rows": int(rows), "of_rows": int(n_local), "ok": ok,
                  "against": "oracle/cgb_oracle.c orc_gather_sum_csr on the same CSR and share rows (the result of the last timed step)"}
        if not ok:
            bad = int((y_gpu != y_cpu).any(axis=1).sum())
            raise RuntimeError(f"parity FAILED: {bad} of {rows} rows of the timed gather differ from the CPU oracle")
        cpu_baseline = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "seconds": secs}
        del y_cpu, y_gpu
        try:  # the scalar port on one thread, on a tenth of the rows (same access pattern)
            v1, s1, _, _, _, _ = cpu_sample(torch, rowptr, col, x, D, 0.1 * args.cpu_frac, reps=2, threads=1, min_seconds=0.0)
            cpu_baseline["value_1_thread"] = v1
            cpu_baseline["sample_1_thread"] = s1
        except Exception as ex:  # noqa: BLE001
            cpu_baseline["value_1_thread_error"] = str(ex)[:200]

    # ---- e2e: host buffers through the C-ABI host entry point, H2D + D2H inside the timed region -----------
    e2e = None
    if not args.no_e2e:
        import ctypes as C

        lib = ctx.lib
        xb = n_local * D * 8
        yb = n_local * D * 8
        hx = torch.empty((n_local, D), dtype=torch.int64).pin_memory()
        hy = torch.empty((n_local, D), dtype=torch.int64).pin_memory()
        hx.copy_(x.cpu())

        def e2e_step():
            if P == 1:  # pipelined host entry point: H2D(i+1) | kernel(i) | D2H(i-1) overlap, all inside the timed region
                ctx.check(lib.cgb_host_gather_sum_async(ctx.handle, csr.handle, C.c_void_p(hx.data_ptr()), None,
                                                        C.c_void_p(hy.data_ptr()), D))
            else:
                x.copy_(hx, non_blocking=True)
                step()
                hy.copy_(v, non_blocking=True)

        def e2e_drain():
            if P == 1:
                ctx.check(lib.cgb_host_sync(ctx.handle))
            else:
                torch.cuda.current_stream().synchronize()

        for _ in range(2):
            e2e_step()
        e2e_drain()
        barrier()
        k2 = max(3, min(K, 20))
        t0 = time.perf_counter()
        for _ in range(k2):
            e2e_step()
        e2e_drain()
        barrier()
        dt = time.perf_counter() - t0
        if P > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": E * P * k2 / dt, "unit": UNIT, "h2d_bytes_per_step": xb, "d2h_bytes_per_step": yb, "steps": k2,
               "api": "cgb_host_gather_sum_async + cgb_host_sync (pinned host share rows in, gathered rows out, every step; "
                      "CSR resident; copies of consecutive steps overlap)" if P == 1
               else "per rank: host share rows -> H2D -> gather + exchange + sum -> D2H of the party's vertex rows"}
        if P == 1:  # the unpipelined call for comparison (one step at a time, copies and kernel serialised)
            t0 = time.perf_counter()
            for _ in range(3):
                ctx.check(lib.cgb_host_gather_sum(ctx.handle, csr.handle, C.c_void_p(hx.data_ptr()), None,
                                                  C.c_void_p(hy.data_ptr()), D))
            e2e["value_single_call_sync"] = E * 3 / (time.perf_counter() - t0)
            e2e["result_equals_device_path"] = bool(torch.equal(hy, y.cpu()))  # every row of the host-buffer result
            assert e2e["result_equals_device_path"]
        else:
            e2e["result_equals_reference_path"] = bool(torch.equal(hy, want_v.cpu()))
            assert e2e["result_equals_reference_path"]
        del hx, hy

    # ---- the hub-free control for the roofline, the matmul record, the epoch record ----------------------------------------
    extra = {}
    if P == 1 and not args.no_uniform:
        try:
            csr.destroy()
            del rowptr, col
            torch.cuda.empty_cache()
            urp, ucol = build_uniform_csr(torch, n_local, E, 99, dev)
            ucsr = ctx.csr_create(urp, ucol, n_local)
            for _ in range(3):
                ctx.gather_sum(ucsr, x, None, out=y)
            ua, ub = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ua.record()
            for _ in range(10):
                ctx.gather_sum(ucsr, x, None, out=y)
            ub.record()
            torch.cuda.synchronize()
            ums = ua.elapsed_time(ub) / 10
            uach = alg / (ums * 1e-3) / 1e9
            roofline["uniform_graph"] = {"graph": "sources and destinations uniform random, same E, N, D (no hubs: a row read is a DRAM read)",
                                         "kernel_ms": ums, "achieved": uach, "unit": "GB/s", "frac": uach / peak,
                                         "edges_per_s": E / (ums * 1e-3)}
            ucsr.destroy()
            del urp, ucol
        except Exception as ex:  # noqa: BLE001
            roofline["uniform_graph"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    if rank == 0 and P == 1 and not args.no_matmul:
        try:
            del y
            torch.cuda.empty_cache()
            extra["matmul"] = matmul_record(torch, ctx, dev)
        except Exception as ex:  # noqa: BLE001
            extra["matmul"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    if not args.no_epoch and not args.no_e2e:
        if P == 1 and rank == 0:
            extra["secure_gcn_epoch"] = secure_gcn_epoch_probe()
        elif P in EPOCH_SHAPE:
            rec = secure_gcn_epoch_nccl(rank, P)
            if rec is not None:
                extra["secure_gcn_epoch"] = rec

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": P, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity}
        if exchange_check:
            line["exchange_check"] = exchange_check
        if phases:
            line["multi_gpu_phases"] = phases
        line.update(extra)
        print(json.dumps(line))
    if P > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
