"""Runs the binaries of oracle/_ref (the reference's own harness / engine / operator code, compiled unchanged against the shim) on
synthetic inputs, one process per party, and parses the reference's log lines.  Used by the CPU test (mock C ABI on the oracle)
and by the GPU test (the real library)."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import run_cluster, synth  # noqa: E402


def graph_dict(edges, tid, feats, labels, cfg):
    """A tests.graphs / synth style graph as the dict synth.write_reference_files takes."""
    import numpy as np

    return {"edges": np.asarray(edges), "tid": np.asarray(tid), "feats": np.asarray(feats, dtype=float), "labels": np.asarray(labels),
            "N": len(tid), "cfg": cfg}


def run(binary, g, T, iters, mock, port_base, timeout=300, extra_env=None):
    exe = os.path.join(ROOT, "oracle", "_ref", binary)
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    d = tempfile.mkdtemp(prefix="cognn_ref_")
    prefix = os.path.join(d, "g")
    synth.write_reference_files(g, prefix)
    cfg = dict(g["cfg"])
    cfg.setdefault("num_samples", len(g["tid"]))
    cfg.setdefault("num_edges", len(g["edges"]))
    cfg.setdefault("test_ratio", 1.0 - cfg["train_ratio"] - cfg["val_ratio"])
    run_cluster.write_config(prefix + "_config.txt", cfg)
    env = dict(os.environ)
    env["COGNN_SHIM_PORT_BASE"] = str(port_base)
    env["OMP_NUM_THREADS"] = "2"
    if mock:
        env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "tests", "mock") + ":" + env.get("LD_LIBRARY_PATH", "")
    env.update(extra_env or {})
    procs, logs = [], []
    for i in range(T):
        cmd = [exe, "-t", str(T), "-g", str(T), "-i", str(i), "-m", str(iters), "-p", "1", "-s", "test", "-c", "0", "-r", "1",
               prefix + ".edge.preprocessed", prefix + ".vertex.preprocessed", prefix + ".part.preprocessed", prefix + ".result",
               prefix + "_config.txt"]
        log = open(os.path.join(d, f"party{i}.log"), "w")
        logs.append(log.name)
        procs.append(subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT, env=env, cwd=d))
    rcs = []
    for p in procs:
        try:
            rcs.append(p.wait(timeout=timeout))
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            rcs.append(-9)
    out = []
    for name in logs:
        txt = open(name).read()
        out.append({"loss": [float(x) for x in re.findall(r"cross-entropy-loss = ([0-9.]+)", txt)],
                    "acc_full": [float(x) for x in re.findall(r"full set accuracy = ([0-9.]+)", txt)],
                    "acc_train": [float(x) for x in re.findall(r"\ntraining set accuracy = ([0-9.]+)", txt)],
                    "acc_test": [float(x) for x in re.findall(r"\ntest set accuracy = ([0-9.]+)", txt)],
                    "iteration_s": [float(x) for x in re.findall(r"::iteration took ([0-9.]+) seconds", txt)],
                    "insecure_banner": "INSECURE" in txt, "tail": txt[-1500:]})
    return rcs, out
