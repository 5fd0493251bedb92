"""Parity at the shapes the numbers are quoted on ("parity at the target shapes"):

  * the arxiv-shaped 8-party graph of BASELINE.json configs[3]: inference (iterations 0-1) and the whole training epoch
    against the epoch oracle -- every share of every party and every message byte;
  * the bench.py headline configuration (RMAT, 100M edges, 6.25M vertices, D = 16, default 128-edge chunks): more than 10^5
    sampled destination rows, the highest-degree rows included, against the C oracle;
  * D = 64 and 128 at 2^24 edges (the size from which the 128-edge chunks are used).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _check_sampled_rows(cgb, oracle, torch, rowptr, col, x, y, n_random, n_top, seed):
    """Compares rows of the device result y with the oracle on a sub-CSR of sampled destination rows."""
    rp = rowptr.cpu().numpy().astype(np.int64)
    deg = np.diff(rp)
    rng = np.random.default_rng(seed)
    rows = np.unique(np.concatenate([rng.integers(0, deg.size, size=n_random), np.argsort(deg)[-n_top:],
                                     np.array([0, deg.size - 1])]))
    sub_deg = deg[rows]
    sub_rp = np.zeros(rows.size + 1, dtype=np.int64)
    sub_rp[1:] = np.cumsum(sub_deg)
    # edge positions of the sampled rows: start of the row + offset inside it
    pos = np.repeat(rp[rows], sub_deg) + (np.arange(int(sub_rp[-1])) - np.repeat(sub_rp[:-1], sub_deg))
    sub_col = col[torch.from_numpy(pos).to(col.device)].cpu().numpy().view(np.uint32)
    xh = x.cpu().numpy().view(np.uint64)
    want = oracle.gather_sum_csr(sub_rp.astype(np.uint32), sub_col, xh)
    got = y[torch.from_numpy(rows).to(y.device)].cpu().numpy().view(np.uint64)
    assert np.array_equal(got, want), f"{int((got != want).any(axis=1).sum())} of {rows.size} sampled rows differ"
    return rows.size, int(sub_deg.max())


def test_bench_headline_configuration_sampled_rows(cgb, oracle):
    import torch

    import bench

    E, D = 100_000_000, 16
    n = E // 16
    rowptr, col = bench.build_party_csr(torch, n, E, 1, 0, 42, torch.device("cuda"))
    csr = cgb.csr_create(rowptr, col, n)
    g = torch.Generator(device="cuda").manual_seed(43)
    x = torch.randint(-2**63, 2**63 - 1, (n, D), dtype=torch.int64, device="cuda", generator=g)
    y = cgb.gather_sum(csr, x)
    assert "256-bit" in cgb.last_kernel
    checked, max_deg = _check_sampled_rows(cgb, oracle, torch, rowptr, col, x, y, 120_000, 64, 1)
    assert checked >= 100_000 and max_deg > 128 * 50  # rows spanning many 128-edge chunks are in the sample
    # checksum of checksums over ALL rows: sum(y) == sum over edges of x[col]
    tot = torch.zeros(D, dtype=torch.int64, device="cuda")
    for lo in range(0, E, 25_000_000):
        tot += x[col[lo:lo + 25_000_000].long()].sum(0)
    assert torch.equal(y.sum(0), tot)
    csr.destroy()


@pytest.mark.parametrize("D", [64, 128])
def test_wide_rows_at_16m_edges_sampled_rows(cgb, oracle, D):
    import torch

    import bench

    E = 1 << 24
    n = E // 16
    rowptr, col = bench.build_party_csr(torch, n, E, 1, 0, 7, torch.device("cuda"))
    csr = cgb.csr_create(rowptr, col, n)
    g = torch.Generator(device="cuda").manual_seed(D)
    x = torch.randint(-2**63, 2**63 - 1, (n, D), dtype=torch.int64, device="cuda", generator=g)
    delta = torch.randint(-2**63, 2**63 - 1, (n, D), dtype=torch.int64, device="cuda", generator=g)
    y = cgb.gather_sum(csr, x)
    checked, _ = _check_sampled_rows(cgb, oracle, torch, rowptr, col, x, y, 20_000, 32, D)
    assert checked >= 15_000
    y2 = cgb.gather_sum(csr, x, delta)
    assert torch.equal(y2, y + delta)
    # the compact block of the same graph, scattered back, is the dense result
    blk = cgb.gather_sum_compact(csr, x)
    v = torch.zeros_like(y)
    cgb.scatter_add_rows(csr.nonempty_rows(), blk, v)
    assert torch.equal(v, y)
    csr.destroy()


def test_arxiv_shaped_8_party_inference_and_epoch_bit_exact():
    """BASELINE configs[3] (and the inference mode of configs[1]) at its own shape: 169 343 vertices, 1.17M edge entries,
    F = 128, H = 16, C = 40, 8 parties, `vid % 8` partition; all parties on this GPU (loopback plane; the NCCL plane runs the
    same engine code and is checked by bench.py --gpus N, `secure_gcn_epoch.bit_exact_vs_oracle`)."""
    from cognn_b200 import engine as eng
    from oracle import epoch as ep
    from tests.test_gpu_engine import NAMES, oracle_tensor
    from tools import synth

    T = 8
    g = synth.make("arxiv", T)
    o = ep.EpochOracle(g["edges"], g["tid"], T, g["feats"], g["labels"], g["cfg"])
    e = eng.Engine(T, g["cfg"], record=True)
    e.load(g["edges"], g["tid"], g["feats"], g["labels"])

    def compare(names):
        for owner in range(T):
            for role in (0, 1):
                for name in names:
                    assert np.array_equal(e.download(owner, role, name), oracle_tensor(o, owner, role, name)), (owner, role, name)

    o.run(2)
    e.run(2)  # inference = iterations 0 and 1 (-m 2, tools/tmp_run_cluster.py:399-415)
    compare(["X", "z0", "z1", "h_t0", "h_t1"])
    o.run(4)
    e.run(4)
    compare(NAMES)
    got = {(m[0], m[1], m[2], m[3]): m[4] for m in e.messages() if not m[3].startswith("setup")}
    want = {(m[0], m[1], m[2], m[3]): m[4] for m in o.msgs}
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    gm = {(m["iter"], m["party"]): m for m in e.metrics()}
    for m in o.log:
        assert abs(gm[(m["iter"], m["party"])]["acc_full"] - m["acc_full"]) < 1e-12
    e.close()
