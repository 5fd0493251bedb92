#!/usr/bin/env python
"""Graph ingest + index-vector construction: device passes (cgb_party_graph_build) against the host builder
(cognn_b200/host/engine.cpp build_party_graph, the reference-shaped restatement of graph_io_util.h:40-208 + ssk.h:295-534).
RMAT graphs of the sweep sizes, `vid % T` partition, one party's tile.  Prints one JSON line per size."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edges", type=int, nargs="+", default=[1_000_000, 10_000_000, 100_000_000])
    ap.add_argument("--parties", type=int, default=8)
    ap.add_argument("--host-max", type=int, default=10_000_000, help="largest size the host builder is timed on")
    args = ap.parse_args()
    import torch

    import bench
    import cognn_b200
    from cognn_b200 import engine as eng

    ctx = cognn_b200.Context(0)
    T = args.parties
    for E in args.edges:
        n = max(T, E // 16)
        src, dst = bench.rmat_edges(torch, n, E, 42, "cuda")
        edges = torch.stack([src, dst], dim=1).contiguous()
        del src, dst
        tid = (torch.arange(n, device="cuda") % T).long()
        rec = {"bench": "party_graph_ingest", "edges": E, "vertices": n, "parties": T, "party": 0}
        for _ in range(2):  # second pass: allocator and CUB temp sizes warm
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g = ctx.party_graph_build(edges, tid, T, 0)
            torch.cuda.synchronize()
            rec["device_resident_input_s"] = time.perf_counter() - t0
            rec["out_edges"] = g["n_out_edges"]
            g["csr"].destroy()
            del g
        he, ht = edges.cpu(), tid.cpu()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = ctx.party_graph_build(he, ht, T, 0)  # host arrays in: H2D of the edge list included
        torch.cuda.synchronize()
        rec["device_host_input_s"] = time.perf_counter() - t0
        g["csr"].destroy()
        del g
        if E <= args.host_max:
            t0 = time.perf_counter()
            w = eng.build_party_graph(he.numpy(), ht.numpy(), T, 0)
            rec["host_builder_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            c = ctx.csr_create(torch.from_numpy(w["rowptr"].view("int32")), torch.from_numpy(w["col"].view("int32")), int(w["vids"].size))
            torch.cuda.synchronize()
            rec["host_builder_plus_csr_upload_s"] = rec["host_builder_s"] + time.perf_counter() - t0
            c.destroy()
            rec["speedup_vs_host"] = rec["host_builder_plus_csr_upload_s"] / rec["device_host_input_s"]
        rec["edges_per_s_device"] = E / rec["device_resident_input_s"]
        print(json.dumps(rec), flush=True)
        del edges, tid
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
