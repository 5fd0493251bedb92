#!/usr/bin/env python
"""Experiment driver in the shape of the reference's tools/tmp_run_cluster.py:run_gcn_test (105-151): writes the three
input files + GNN config of a synthetic named shape, spawns one `gcn-optimize-b200` process per party (one GPU each,
NCCL between them instead of netns + TCP), and parses the reference's log lines (`::iteration took`, accuracy)."""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import synth  # noqa: E402


def write_config(path, cfg):
    with open(path, "w") as f:
        f.write("num_layers : 2\n" + "\n".join(f"{k} : {cfg[k]}" for k in ("num_labels", "input_dim", "hidden_dim", "num_samples", "num_edges",
                                                                          "learning_rate", "train_ratio", "val_ratio", "test_ratio")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="cora")
    ap.add_argument("--parties", type=int, default=2)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--inter", type=float, default=None)
    ap.add_argument("--loopback", action="store_true", help="all parties in one process on GPU 0")
    args = ap.parse_args()
    from cognn_b200.host import build as hb

    exe = hb.build_harness()
    T = args.parties
    g = synth.make(args.shape, T, args.inter)
    d = tempfile.mkdtemp(prefix="cognn_b200_")
    prefix = os.path.join(d, args.shape)
    synth.write_reference_files(g, prefix)
    write_config(prefix + "_config.txt", g["cfg"])
    procs, logs = [], []
    parties = [0] if args.loopback else list(range(T))
    for i in parties:
        cmd = [exe, "-t", str(T), "-g", str(T), "-i", str(i), "-m", str(args.iters), "-p", "1", "-s", f"gcn-optimize/{args.shape}/{T}p",
               "-r", "1", prefix + ".edge.preprocessed", prefix + ".vertex.preprocessed", prefix + ".part.preprocessed",
               prefix + ".result", prefix + "_config.txt"]
        env = dict(os.environ)
        env.setdefault("COGNN_B200_ALLOW_INSECURE_EMULATION", "1")  # a benchmark driver: dealer emulation acknowledged
        env.setdefault("COGNN_B200_KEY", "2d,0,0,0,0,0,0,0")
        if args.loopback:
            env["COGNN_B200_PLANE"] = "loopback"
        else:
            env["COGNN_B200_DEVICE"] = str(i)
        log = open(os.path.join(d, f"gcn_test_{args.shape}_{i}.log"), "w")
        logs.append(log.name)
        procs.append(subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT, env=env))
    rcs = [p.wait() for p in procs]
    out = {"bench": "run_cluster", "shape": args.shape, "parties": T, "iters": args.iters, "plane": "loopback" if args.loopback else "nccl",
           "return_codes": rcs, "per_party": []}
    for name in logs:
        txt = open(name).read()
        it = [float(x) for x in re.findall(r"::iteration took ([0-9.]+) seconds", txt)]
        acc = [float(x) for x in re.findall(r"full set accuracy = ([0-9.]+)", txt)]
        loss = [float(x) for x in re.findall(r"cross-entropy-loss = ([0-9.]+)", txt)]
        out["per_party"].append({"iteration_s": it, "epoch_s_last": sum(it[-6:]) if len(it) >= 6 else sum(it), "full_set_accuracy": acc,
                                 "loss": loss, "tail": txt[-300:] if any(rcs) else ""})
    print(json.dumps(out))
    return 1 if any(rcs) else 0


if __name__ == "__main__":
    sys.exit(main())
