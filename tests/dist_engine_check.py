#!/usr/bin/env python
"""Launched with torchrun on T GPUs (one party per GPU): the engine over its NCCL plane must produce exactly the
oracle's shares and messages.  Not collected by pytest (run by gpurun --gpus N); prints one JSON line on rank 0.

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_engine_check.py
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cognn_b200 import engine as eng  # noqa: E402
from oracle import epoch as ep  # noqa: E402
from tests.graphs import small_graph  # noqa: E402
from tests.test_gpu_engine import NAMES, oracle_tensor  # noqa: E402


def main():
    rank, T = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (ctypes.c_char * 128)()
        assert eng.load_host().cge_nccl_unique_id(buf) == 0
        uid = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
    dist.broadcast(uid, 0)
    uid2 = torch.zeros(128, dtype=torch.uint8)  # a second NCCL communicator for the graph-replay leg
    if rank == 0:
        buf = (ctypes.c_char * 128)()
        assert eng.load_host().cge_nccl_unique_id(buf) == 0
        uid2 = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
    dist.broadcast(uid2, 0)
    g = small_graph(n=90, n_edges=400, F=12, C=4, T=T, seed=77)
    cfg = dict(input_dim=12, hidden_dim=8, num_labels=4, learning_rate=0.5, train_ratio=0.4, val_ratio=0.2)
    e = eng.Engine(T, cfg, device=local_rank, rank=rank, nccl_uid=uid.numpy().tobytes(), record=True)
    e.load(g["edges"], g["tid"], g["feats"], g["labels"])
    e.run(12)
    o = ep.EpochOracle(g["edges"], g["tid"], T, g["feats"], g["labels"], cfg)
    o.run(12)
    bad = []
    # this rank hosts: share 0 of owner `rank`, share 1 of owner rank - 1
    for owner, role in ((rank, 0), ((rank - 1) % T, 1)):
        for name in NAMES:
            if not np.array_equal(e.download(owner, role, name), oracle_tensor(o, owner, role, name)):
                bad.append((owner, role, name))
    got = {(m[0], m[1], m[2], m[3]): m[4] for m in e.messages() if not m[3].startswith("setup")}
    want = {(m[0], m[1], m[2], m[3]): m[4] for m in o.msgs if m[1] == rank}
    if set(got) != set(want):
        bad.append(("message keys", len(got), len(want)))
    else:
        bad += [k for k in want if not np.array_equal(got[k], want[k])]
    e.close()
    # second leg: no transcript recorder, so from epoch 1 on each iteration's online phase (NCCL send/recv groups included)
    # is captured into a CUDA graph and replayed in epochs 2 and 3; the run ends in the middle of an epoch
    n2 = 20
    e = eng.Engine(T, cfg, device=local_rank, rank=rank, nccl_uid=uid2.numpy().tobytes())
    e.load(g["edges"], g["tid"], g["feats"], g["labels"])
    e.run(n2)
    replays = e.graph_replays
    o = ep.EpochOracle(g["edges"], g["tid"], T, g["feats"], g["labels"], cfg)
    o.run(n2)
    for owner, role in ((rank, 0), ((rank - 1) % T, 1)):
        for name in NAMES:
            if not np.array_equal(e.download(owner, role, name), oracle_tensor(o, owner, role, name)):
                bad.append(("graph", owner, role, name))
    if os.environ.get("COGNN_B200_GRAPHS", "1") != "0" and replays != n2 - 12:
        bad.append(("graph replays", replays))
    if bad:
        sys.stderr.write(f"[rank {rank}] mismatches: {bad[:40]}\n")
    res = torch.tensor([len(bad)], dtype=torch.int64)
    dist.all_reduce(res)
    if rank == 0:
        print(json.dumps({"check": "engine_nccl_vs_oracle", "parties": T, "iterations": [12, n2], "mismatches": int(res.item()),
                          "messages_rank0": len(got), "graph_replays_rank0": replays, "ok": int(res.item()) == 0}), flush=True)
    e.close()
    dist.destroy_process_group()
    sys.exit(0 if int(res.item()) == 0 else 1)


if __name__ == "__main__":
    main()
