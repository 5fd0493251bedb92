"""The CLI drop-in (cognn_b200/host/harness.cpp): the reference's flags, the reference's three input files and config
file, the reference's log lines.  Runs all parties in one process (COGNN_B200_PLANE=loopback) on the cora_small-sized
fixture and on a Cora-shaped graph, and checks the printed accuracies against the engine driven through its C ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_config(path, cfg):
    with open(path, "w") as f:
        f.write(f"num_layers : 2\nnum_labels : {cfg['num_labels']}\ninput_dim : {cfg['input_dim']}\nhidden_dim : {cfg['hidden_dim']}\n"
                f"num_samples : {cfg['num_samples']}\nnum_edges : {cfg['num_edges']}\nlearning_rate : {cfg['learning_rate']}\n"
                f"train_ratio : {cfg['train_ratio']}\nval_ratio : {cfg['val_ratio']}\ntest_ratio : {cfg['test_ratio']}")


def test_cli_rejects_bad_arguments_like_the_reference():
    from cognn_b200.host import build as hb

    exe = hb.build_harness()
    r = subprocess.run([exe, "-i", "0"], capture_output=True, text=True)
    assert r.returncode != 0 and "Must specify number of threads and number of graph tiles." in r.stderr
    r = subprocess.run([exe, "-t", "2", "-g", "3", "x"], capture_output=True, text=True)
    assert r.returncode != 0 and "Number of threads must be a divisor of number of graph tiles." in r.stderr
    r = subprocess.run([exe, "-t", "2", "-g", "2"], capture_output=True, text=True)
    assert r.returncode != 0 and "Must specify an input edge list file." in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("shape,T,iters", [("cora_small", 2, 6), ("cora", 2, 12)])
def test_cli_runs_reference_inputs(tmp_path, shape, T, iters):
    from cognn_b200 import engine as eng
    from cognn_b200.host import build as hb
    from tools import synth

    g = synth.make(shape, T)
    prefix = str(tmp_path / shape)
    synth.write_reference_files(g, prefix)
    write_config(prefix + "_config.txt", g["cfg"])
    exe = hb.build_harness()
    env = dict(os.environ, COGNN_B200_PLANE="loopback")
    # the binary refuses to run until the caller acknowledges the dealer emulation and supplies the master key
    r = subprocess.run([exe, "-t", str(T), "-g", str(T), "-i", "0", "-m", "1", "-r", "1", prefix + ".edge.preprocessed",
                        prefix + ".vertex.preprocessed", prefix + ".part.preprocessed", prefix + ".result", prefix + "_config.txt"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode != 0 and "INSECURE" in r.stdout and "refusing to run" in r.stdout
    env.update(COGNN_B200_ALLOW_INSECURE_EMULATION="1", COGNN_B200_KEY="2d,0,0,0,0,0,0,0")  # the key the Engine wrapper defaults to
    cmd = [exe, "-t", str(T), "-g", str(T), "-i", "0", "-m", str(iters), "-p", "1", "-s", "gcn-optimize/x/2p", "-c", "0", "-r", "1",
           prefix + ".edge.preprocessed", prefix + ".vertex.preprocessed", prefix + ".part.preprocessed", prefix + ".result",
           prefix + "_config.txt"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert out.count("::iteration took") == iters and "::preprocess took" in out and "Finish algo kernel" in out and "INSECURE" in out
    acc = [float(x) for x in re.findall(r"full set accuracy = ([0-9.]+)", out)]
    e = eng.Engine(T, g["cfg"])
    e.load(g["edges"], g["tid"], g["feats"], g["labels"])
    e.run(iters)
    want = [m["acc_full"] for m in e.metrics()]
    assert len(acc) == len(want) and np.allclose(acc, want, atol=1e-6)
    assert os.path.exists(prefix + ".result")
    e.close()
