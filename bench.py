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
parties by `vid % N` (tools/data_transform.py:25): the step is the fused gather into one N_p x D block per
destination party, the NCCL all-to-all of those mirror-update blocks (ssk.h:835 -> 1067/1090) and
the share-local sum of the received blocks (GatherComp add, optimize-gcn/gcn.h:456).  Weak scaling: edges per GPU
are fixed.  WAN / 2PC-residual costs are out of scope on both arms.

The reference cannot be built here (its arithmetic is in un-vendored trees), so `--impl reference` and
`cpu_baseline` time the CPU oracle (oracle/cgb_oracle.c, a restatement of the same path; kind = "port") with all
host threads on a bounded sample of the same graph.
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


def build_party_csr(torch, n_local, n_edges, n_parties, rank, seed, device):
    """CSR-by-destination of the edges party `rank` owns.  Sources are its local vertices (row index 0..n_local),
    destinations are global vertices grouped by owner: output row of global vertex v = (v % P) * n_local + v // P."""
    n_global = n_local * n_parties
    src, dst = rmat_edges(torch, n_global, n_edges, seed + 1000 * rank, device)
    src = src % n_local  # the party's own vertices (local row index)
    out_row = (dst % n_parties) * n_local + dst // n_parties
    del dst
    key = out_row * n_local + src  # sort by destination row, sources ascending inside a row (ssk.h:295-314 order)
    del out_row, src
    key = torch.sort(key).values
    out_row = key // n_local
    col = (key - out_row * n_local).int()
    del key
    counts = torch.bincount(out_row, minlength=n_global)
    del out_row
    rowptr = torch.zeros(n_global + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr.int(), col


def exchange_and_sum(dist, y, recv, v, n_parties, n_local, D, add):
    """Mirror-update exchange of one GAS iteration (ssk.h:835 -> 1067/1090): block i of y goes to party i, then the
    received blocks are summed share-locally (GatherComp add, gcn.h:456).  `add(a, b, out)` is the engine's add."""
    dist.all_to_all_single(recv, y)
    blocks = recv.view(n_parties, n_local, D)
    add(blocks[0], blocks[1], v)
    for j in range(2, n_parties):
        add(v, blocks[j], v)
    return v


class RawCuda:
    """A cgb_malloc allocation viewed as a torch int64 tensor (cudaMalloc memory is IPC-exportable, torch's is not)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def setup_peer_windows(torch, dist, ctx, rank, P, n_local, D, dev):
    """Double-buffered receive windows, mapped into every peer with CUDA IPC.  Window k of party t holds one n_local x D
    block per source party; the gather kernel of party `rank` stores its block for t straight into it over NVLink."""
    nbytes = P * n_local * D * 8
    bufs = [ctx.malloc(nbytes) for _ in range(2)]
    mine = [ctx.ipc_export(b) for b in bufs]
    everyone = [None] * P
    dist.all_gather_object(everyone, mine)
    block_ptrs = []
    for k in range(2):
        row = []
        for t in range(P):
            base = bufs[k] if t == rank else ctx.ipc_open(everyone[t][k])
            row.append(base + rank * n_local * D * 8)
        block_ptrs.append(row)
    views = [torch.as_tensor(RawCuda(b, (P, n_local, D)), device=dev) for b in bufs]
    offsets = [t * n_local for t in range(P + 1)]
    return {"bufs": bufs, "block_ptrs": block_ptrs, "views": views, "offsets": offsets,
            "flag": torch.zeros(1, dtype=torch.int32, device=dev)}


def fused_step(dist, ctx, csr, x, win, step_idx, v, P, add):
    """One GAS gather step with the exchange fused into the kernel: every output row is written once, directly into the
    window of the party that owns it (peer memory over NVLink); a 4-byte all-reduce is the cross-rank barrier."""
    k = step_idx & 1
    ctx.gather_sum_blocks(csr, x, win["block_ptrs"][k], win["offsets"])
    dist.all_reduce(win["flag"])  # stream-ordered barrier: all peers have finished writing window k
    blocks = win["views"][k]
    ctx.sum_n([blocks[j] for j in range(P)], out=v)  # GatherComp additions over all source parties, one pass
    return v


def setup_pipelined(torch, ctx, win, rowptr, col, rank, P, n_local, D, dev):
    """One CSR per destination party, a local staging block per remote party and views of the peers' windows: the
    gather of block t+1 overlaps the NVLink copy (DMA) of block t."""
    csrs = []
    for t in range(P):
        lo, hi = t * n_local, (t + 1) * n_local
        e0, e1 = int(rowptr[lo]), int(rowptr[hi])
        csrs.append(ctx.csr_create((rowptr[lo:hi + 1] - rowptr[lo]).int().contiguous(), col[e0:e1].contiguous(), n_local))
    stage = torch.empty((P, n_local, D), dtype=torch.int64, device=dev)
    slot_bytes = n_local * D * 8
    peer = [[torch.as_tensor(RawCuda(win["block_ptrs"][k][t], (n_local, D)), device=dev) for t in range(P)] for k in range(2)]
    copy_streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(P)]  # high priority: copy CTAs are placed ahead of the next gather's
    copy_ctx = []
    import cognn_b200
    for t in range(P):  # one context per copy stream: cgb_peer_copy enqueues on its context's stream
        with torch.cuda.stream(copy_streams[t]):
            copy_ctx.append(cognn_b200.Context(dev.index))
    return {"csrs": csrs, "stage": stage, "peer": peer, "copy_streams": copy_streams, "copy_ctx": copy_ctx,
            "ev": [torch.cuda.Event(enable_timing=True) for _ in range(P)],
            "done": [torch.cuda.Event(enable_timing=True) for _ in range(P)],
            "t0": torch.cuda.Event(enable_timing=True), "bar": torch.cuda.Event(enable_timing=True),
            "end": torch.cuda.Event(enable_timing=True),
            "slot_bytes": slot_bytes,
            # transport of the pushes (profiles/r1_p2p_probe_n2.jsonl, r1b_bench_n{2,4,8}*): the SM copy kernel moves 690 GB/s
            # per GPU against 537 GB/s for the copy engines; it takes SM slots from the gathers beside it, but wins at every
            # party count measured (N=2: 2.99 vs 3.14 ms, N=4: 5.5 vs 6.1 ms).  CGB_BENCH_COPY=ce selects the copy engines.
            "copy": os.environ.get("CGB_BENCH_COPY", "sm"),
            "copy_ctas": int(os.environ.get("CGB_BENCH_COPY_CTAS", "64"))}


def pipelined_step(torch, dist, ctx, x, win, pipe, step_idx, v, rank, P):
    """Gather per destination block (remote blocks first), push each finished block to its owner with an SM-driven peer copy
    (cgb_peer_copy, 64 CTAs of 128 threads) on a side stream while the next block is gathered; 4-byte all-reduce as barrier; one-pass sum of
    the received blocks."""
    k = step_idx & 1
    main = torch.cuda.current_stream()
    pipe["t0"].record(main)
    for j in range(1, P + 1):
        t = (rank + j) % P
        if t != rank:
            ctx.gather_sum(pipe["csrs"][t], x, None, out=pipe["stage"][t])
            pipe["ev"][t].record(main)
            cs = pipe["copy_streams"][t]  # one copy stream per destination: copies to different peers run concurrently
            cs.wait_event(pipe["ev"][t])
            if pipe["copy"] == "sm":  # SM-driven push: 64 small CTAs saturate the NVLink egress (profiles/r1_p2p_probe_n2.jsonl)
                pipe["copy_ctx"][t].peer_copy(win["block_ptrs"][k][t], pipe["stage"][t].data_ptr(), pipe["slot_bytes"],
                                              pipe["copy_ctas"])
            else:  # copy engine (cudaMemcpyAsync peer copy)
                with torch.cuda.stream(cs):
                    pipe["peer"][k][t].copy_(pipe["stage"][t], non_blocking=True)
            pipe["done"][t].record(cs)
        else:
            ctx.gather_sum(pipe["csrs"][t], x, None, out=pipe["peer"][k][t])  # own window, own slot
            pipe["ev"][t].record(main)
    for t in range(P):
        if t != rank:
            main.wait_event(pipe["done"][t])
    dist.all_reduce(win["flag"])
    pipe["bar"].record(main)
    blocks = win["views"][k]
    ctx.sum_n([blocks[j] for j in range(P)], out=v)
    pipe["end"].record(main)
    return v


def pipelined_phases(pipe, rank, P):
    """Timeline of the LAST pipelined step on this rank (ms after the step's start), read after a synchronize."""
    t0 = pipe["t0"]
    order = [(rank + j) % P for j in range(1, P + 1)]
    return {"gather_done": [round(t0.elapsed_time(pipe["ev"][t]), 3) for t in order],
            "copy_done": [round(t0.elapsed_time(pipe["done"][t]), 3) for t in order if t != rank],
            "barrier_done": round(t0.elapsed_time(pipe["bar"]), 3), "sum_done": round(t0.elapsed_time(pipe["end"]), 3)}


def secure_gcn_epoch_probe(timeout_s=240):
    """BASELINE.json configs[0]: 2-party CoGNN-Opt training epoch on the synthetic Cora-shaped graph, both parties on this GPU
    (loopback plane; `tools/epoch_bench.py` under torchrun is the one-party-per-GPU form).  Runs in a child process so that
    nothing it does can disturb the gather measurement; any failure is reported, never raised."""
    import subprocess

    cmd = [sys.executable, os.path.join(ROOT, "tools", "epoch_bench.py"), "--shape", "cora", "--parties", "2", "--epochs", "4", "--cpu"]
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        rec = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        keep = ("shape", "parties", "plane", "N", "E", "cfg", "iterations", "online_s", "online_mode", "offline_dealer_s", "launches",
                "rounds", "graph_replays", "cpu_oracle", "note")
        out = {k: rec[k] for k in keep if k in rec}
        out["unit"] = "s per epoch (online phase; offline = trusted-dealer emulation, reported beside it)"
        return out
    except Exception as ex:  # noqa: BLE001
        return {"error": f"{type(ex).__name__}: {ex}"[:300]}


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


def ncu_traffic(D):
    """dram bytes per gather_sum launch from the committed ncu --set full capture of this command, or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("gather_sum_kernel", {}).get(f"D{D}_bytes_per_launch")
        except Exception:
            return None
    return None


def cpu_sample(torch, rowptr, col, x, D, frac_rows, reps=3, threads=None, min_seconds=10.0):
    """Times the CPU oracle on the first `frac_rows` destination rows of the SAME graph (same index skew, same
    random row reads into the full x).  Returns (edges/s, description, threads)."""
    import numpy as np

    from oracle import pyoracle

    pyoracle.build()
    if threads:
        pyoracle.set_num_threads(threads)
    n_rows = rowptr.numel() - 1
    rows = max(1, int(n_rows * frac_rows))
    rp = rowptr[: rows + 1].cpu().numpy().view(np.uint32).copy()
    e = int(rp[-1])
    cl = col[:e].cpu().numpy().view(np.uint32).copy()
    xh = x.cpu().numpy().view(np.uint64)
    best, total, n = None, 0.0, 0
    while n < reps or (total < min_seconds and n < 200):  # about 10 s of CPU work
        t0 = time.perf_counter()
        pyoracle.gather_sum_csr(rp, cl, xh)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        total += dt
        n += 1
    return e / best, (f"first {rows} of {n_rows} destination rows ({e} edges) of the same graph, best of {n} passes "
                      f"({total:.1f} s of CPU work)"), pyoracle.num_threads(), best


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
    ap.add_argument("--no-epoch", action="store_true", help="skip the secure-GCN epoch probe (N = 1 only)")
    ap.add_argument("--exchange", default="pipelined", choices=["pipelined", "fused", "nccl"],
                    help="N > 1: pipelined = per-destination gathers overlapped with async peer copies over NVLink (default); "
                         "fused = one gather kernel storing straight into peer windows; nccl = gather then all_to_all")
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
        "exchange": {"pipelined": "one gather per destination party; each finished block is pushed into its owner's window "
                                  "(CUDA IPC peer memory over NVLink) by an SM-driven copy kernel that overlaps the next gather; "
                                  "4-byte all-reduce as barrier; one-pass sum",
                     "fused": "one gather kernel stores each block straight into the consumer's window over NVLink; "
                              "4-byte all-reduce as barrier",
                     "nccl": "gather, then NCCL all_to_all_single of the blocks"}[args.exchange] if P > 1
        else "none (single party)",
    }

    # -------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        if dev == "cpu":
            E_ref = min(E, 2_000_000)  # local smoke only
            n_ref = max(1, E_ref // 16)
        else:
            E_ref, n_ref = E, n_local
        rowptr, col = build_party_csr(torch, n_ref, E_ref, 1, 0, 42, dev)
        g = torch.Generator(device=dev).manual_seed(43)
        x = torch.randint(-2**63, 2**63 - 1, (n_ref, D), dtype=torch.int64, device=dev, generator=g)
        import numpy as np

        from oracle import pyoracle

        pyoracle.build()
        rows = max(1, int(n_ref * args.cpu_frac))
        rp = rowptr[: rows + 1].cpu().numpy().view(np.uint32).copy()
        e = int(rp[-1])
        cl = col[:e].cpu().numpy().view(np.uint32).copy()
        xh = x.cpu().numpy().view(np.uint64)
        del rowptr, col, x
        for _ in range(args.warmup):
            pyoracle.gather_sum_csr(rp, cl, xh)
        t0 = time.perf_counter()
        for _ in range(K):
            pyoracle.gather_sum_csr(rp, cl, xh)
        dt = time.perf_counter() - t0
        val = e * K / dt
        sample = f"each step = first {rows} of {n_ref} destination rows ({e} edges) of the same graph"
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
                "warmup": args.warmup, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": pyoracle.num_threads(), "kind": "port",
                                 "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "CPU oracle port of the reference path (reference unbuildable here, SURVEY.md 8c); host cores only"}
        print(json.dumps(line))
        return 0

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
    csr = ctx.csr_create(rowptr, col, n_local)
    g = torch.Generator(device=dev).manual_seed(43 + rank)
    x = torch.randint(-2**63, 2**63 - 1, (n_local, D), dtype=torch.int64, device=dev, generator=g)
    y = torch.empty((n_local * P, D), dtype=torch.int64, device=dev)
    recv = torch.empty_like(y) if P > 1 else None
    v = torch.empty((n_local, D), dtype=torch.int64, device=dev) if P > 1 else None

    add = lambda a, b, o: ctx.add(a, b, out=o)  # noqa: E731
    fused = P > 1 and args.exchange == "fused"
    piped = P > 1 and args.exchange == "pipelined"
    win = pipe = None
    if fused or piped:
        # peer windows need CUDA IPC between the ranks' processes; if the platform refuses it, say so and use the NCCL
        # all-to-all exchange (same kernels, same result) instead of failing the whole run
        try:
            win = setup_peer_windows(torch, dist, ctx, rank, P, n_local, D, dev)
            ok = torch.ones(1, dtype=torch.int32, device=dev)
        except Exception as ex:  # noqa: BLE001
            sys.stderr.write(f"[bench] rank {rank}: peer windows unavailable ({ex}); falling back to --exchange nccl\n")
            ok = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            fused = piped = False
            win = None
            config["exchange"] = "gather, then NCCL all_to_all_single of the blocks (peer windows unavailable on this platform)"
    if piped:
        pipe = setup_pipelined(torch, ctx, win, rowptr, col, rank, P, n_local, D, dev)
    step_no = [0]
    if fused:
        # self-check outside the timed region: the fused path must equal gather + NCCL all-to-all + sum
        ctx.gather_sum(csr, x, None, out=y)
        ref = exchange_and_sum(dist, y, recv, torch.empty_like(v), P, n_local, D, add).clone()
        for k in range(2):
            got = fused_step(dist, ctx, csr, x, win, k, v, P, add)
            assert torch.equal(got, ref), "fused peer-store exchange differs from the NCCL all-to-all path"
        dist.barrier()
    if piped:
        ctx.gather_sum(csr, x, None, out=y)
        ref = exchange_and_sum(dist, y, recv, torch.empty_like(v), P, n_local, D, add).clone()
        for k in range(2):
            got = pipelined_step(torch, dist, ctx, x, win, pipe, k, v, rank, P)
            assert torch.equal(got, ref), "pipelined peer-copy exchange differs from the NCCL all-to-all path"
        dist.barrier()

    def step():
        if piped:
            pipelined_step(torch, dist, ctx, x, win, pipe, step_no[0], v, rank, P)
            step_no[0] += 1
            return
        if fused:
            fused_step(dist, ctx, csr, x, win, step_no[0], v, P, add)
            step_no[0] += 1
            return
        ctx.gather_sum(csr, x, None, out=y)
        if P > 1:
            exchange_and_sum(dist, y, recv, v, P, n_local, D, add)

    def barrier():
        if P > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    # kernel-only duration of the dominant kernel (gather_sum) for the roofline, CUDA events on the launch stream
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    bev = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    sev = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    all_ctx = [ctx] + (pipe["copy_ctx"] if pipe else [])
    launches0 = sum(c.launches for c in all_ctx)
    t_wall0 = time.time()
    torch.cuda.profiler.start()  # ncu --profile-from-start off captures exactly the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        if P == 1:
            kev[i][0].record()
            step()
            kev[i][1].record()
        elif piped:
            kev[i][0].record()
            step()
            kev[i][1].record()
        elif fused:
            k = step_no[0] & 1
            kev[i][0].record()
            ctx.gather_sum_blocks(csr, x, win["block_ptrs"][k], win["offsets"])
            kev[i][1].record()
            dist.all_reduce(win["flag"])
            bev[i].record()
            blocks = win["views"][k]
            ctx.sum_n([blocks[j] for j in range(P)], out=v)
            sev[i].record()
            step_no[0] += 1
        else:
            kev[i][0].record()
            ctx.gather_sum(csr, x, None, out=y)
            kev[i][1].record()
            exchange_and_sum(dist, y, recv, v, P, n_local, D, add)
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    t_wall1 = time.time()
    launches = sum(c.launches for c in all_ctx) - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / K
    piped_timeline = None
    if piped:
        allp = [None] * P
        mine = pipelined_phases(pipe, rank, P)
        mine["edges_per_block_in_gather_order"] = [pipe["csrs"][(rank + j) % P].n_edges for j in range(1, P + 1)]
        dist.all_gather_object(allp, mine)
        piped_timeline = allp
    if piped:
        # the P per-destination gather launches of one step, timed on their own (no copies) for the roofline
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        ka.record()
        for _ in range(reps):
            for t in range(P):
                ctx.gather_sum(pipe["csrs"][t], x, None, out=pipe["stage"][t])
        kb.record()
        torch.cuda.synchronize()
        kernel_ms = ka.elapsed_time(kb) / reps
    if P > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = E * P / (ms_per_step * 1e-3)
    phases = None
    if fused:
        mine = torch.tensor([kernel_ms, sum(kev[i][1].elapsed_time(bev[i]) for i in range(K)) / K,
                             sum(bev[i].elapsed_time(sev[i]) for i in range(K)) / K], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(P)]
        dist.all_gather(allr, mine)
        phases = {"per_rank_ms": {"gather_kernel_with_peer_stores": [round(float(t[0]), 3) for t in allr],
                                  "barrier_wait": [round(float(t[1]), 3) for t in allr],
                                  "sum_of_received_blocks": [round(float(t[2]), 3) for t in allr]}}

    peak, peak_src = measured_peak_hbm()
    n_rows = n_local * P
    alg = algorithmic_bytes(n_rows, E, D)
    achieved = alg / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "gather_chunk_kernel<4,4,4,128,1024,2> (cgb_gather_sum, 256-bit row loads)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(D), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "kernel_ms": kernel_ms,
                "frac_of_8TBs_nominal": achieved / 8000.0,
                "note": "algorithmic bytes charge every edge a full 8*D-byte row read (SURVEY 8d, no cache-reuse credit); hub rows "
                        "are served by L2, so `traffic` (DRAM bytes per launch, ncu) is below them and frac can exceed 1"}

    # ---- e2e: host buffers through the C-ABI host entry point, H2D + D2H inside the timed region -----------
    e2e = None
    if not args.no_e2e:
        import ctypes as C

        lib = ctx.lib
        xb, yb = n_local * D * 8, n_rows * D * 8
        hx = torch.empty((n_local, D), dtype=torch.int64).pin_memory()
        hy = torch.empty((n_rows, D), dtype=torch.int64).pin_memory()
        hx.copy_(x.cpu())
        hv = torch.empty((n_local, D), dtype=torch.int64).pin_memory() if P > 1 else None

        def e2e_step():
            if P == 1:  # pipelined host entry point: H2D(i+1) | kernel(i) | D2H(i-1) overlap, all inside the timed region
                ctx.check(lib.cgb_host_gather_sum_async(ctx.handle, csr.handle, C.c_void_p(hx.data_ptr()), None,
                                                        C.c_void_p(hy.data_ptr()), D))
            else:
                x.copy_(hx, non_blocking=True)
                step()
                hv.copy_(v, non_blocking=True)

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
        e2e = {"value": E * P * k2 / dt, "unit": UNIT, "h2d_bytes_per_step": xb,
               "d2h_bytes_per_step": yb if P == 1 else n_local * D * 8, "steps": k2,
               "api": "cgb_host_gather_sum_async + cgb_host_sync (pinned host share rows in, gathered rows out, every step; "
                      "CSR resident; copies of consecutive steps overlap)" if P == 1
               else "host share rows -> H2D -> gather + exchange + sum -> D2H"}
        if P == 1:  # the unpipelined call for comparison (one step at a time, copies and kernel serialised)
            t0 = time.perf_counter()
            for _ in range(3):
                ctx.check(lib.cgb_host_gather_sum(ctx.handle, csr.handle, C.c_void_p(hx.data_ptr()), None,
                                                  C.c_void_p(hy.data_ptr()), D))
            e2e["value_single_call_sync"] = E * 3 / (time.perf_counter() - t0)
        # spot check of the e2e result against the device-resident result
        if P == 1:
            assert torch.equal(hy[:1000], y[:1000].cpu())

    # ---- CPU baseline on rank 0 at N == 1 -------------------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and P == 1 and not args.no_cpu_baseline:
        val, sample, cores, secs = cpu_sample(torch, rowptr, col, x, D, args.cpu_frac)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                        "seconds": secs}
        try:  # the scalar port on one thread, on a tenth of the rows (same access pattern)
            v1, s1, _, _ = cpu_sample(torch, rowptr, col, x, D, 0.1, reps=2, threads=1, min_seconds=0.0)
            cpu_baseline["value_1_thread"] = v1
            cpu_baseline["sample_1_thread"] = s1
        finally:
            from oracle import pyoracle as _po

            _po.set_num_threads(cores)

    # ---- the other half of BASELINE.json's metric: secure-GCN epoch time (configs[0] shape), in a child process ------------
    epoch = None
    if rank == 0 and P == 1 and not args.no_epoch and not args.no_e2e:
        epoch = secure_gcn_epoch_probe()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": P, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline}
        if phases:
            line["multi_gpu_phases"] = phases
        if piped_timeline:
            line["multi_gpu_phases"] = {"per_rank_last_step_timeline_ms": piped_timeline}
        if epoch is not None:
            line["secure_gcn_epoch"] = epoch
        print(json.dumps(line))
    if P > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
