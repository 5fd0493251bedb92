"""(COGNN_B200_PEER_EXCHANGE=1 forces the peer plane, which is on by default only above 4 parties.)
Tests that need two GPUs on the box (skipped otherwise): the engine's epoch with one party per GPU, protocol rounds over the
peer-memory plane (cgb_peer_round through CUDA-IPC mappings) and over ncclSend / ncclRecv, both bit exact against the epoch
oracle.  One process per GPU under torchrun, rendezvous on 127.0.0.1."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("peer", ["1", "0"])
def test_engine_epoch_two_gpus_bit_exact(peer):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, COGNN_B200_PEER_EXCHANGE=peer, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "tools", "epoch_bench.py"), "--shape", "cora",
           "--parties", "2", "--epochs", "3", "--check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    rec = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert rec["bit_exact_vs_oracle"] is True and rec["checked"]["mismatches"] == 0
    assert rec["plane"].startswith("peer-memory" if peer == "1" else "nccl"), rec["plane"]
    assert rec["graph_replays"] > 0 and rec["rounds"] == 27
