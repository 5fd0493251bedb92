"""GPU parity (bit exact) of op (1) -- gather-sum / expand / segsum -- against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

from tests.util import power_law_csr, rand_u64, to_dev, to_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D", [1, 2, 3, 7, 16, 40, 64, 100, 128, 131])
def test_gather_sum_matches_oracle(cgb, oracle, D):
    rng = np.random.default_rng(100 + D)
    n_dst, n_src = 3000, 2500
    rowptr, col = power_law_csr(rng, n_dst, n_src, 40000)
    assert (np.diff(rowptr.astype(np.int64)) > 256).any(), "case must contain long (sliced) rows"
    x, delta = rand_u64(rng, n_src, D), rand_u64(rng, n_dst, D)
    csr = cgb.csr_create(to_dev(rowptr, "cpu"), to_dev(col, "cpu"), n_src)
    dx, dd = to_dev(x), to_dev(delta)
    for use_delta in (False, True):
        want = oracle.gather_sum_csr(rowptr, col, x, delta if use_delta else None)
        got = cgb.gather_sum(csr, dx, dd if use_delta else None)
        assert np.array_equal(to_np(got), want)
        # second launch reuses the self-resetting arrival counters
        got = cgb.gather_sum(csr, dx, dd if use_delta else None)
        assert np.array_equal(to_np(got), want)
    csr.destroy()


def test_gather_sum_device_csr_and_edge_cases(cgb, oracle):
    import torch

    rng = np.random.default_rng(7)
    # empty graph rows, single row, all edges in one row (one very long row), zero rows
    for n_dst, degs in [(5, [0, 0, 0, 0, 0]), (1, [1]), (3, [0, 5000, 0]), (4, [257, 256, 255, 1])]:
        rowptr = np.zeros(n_dst + 1, dtype=np.uint32)
        rowptr[1:] = np.cumsum(degs)
        n_src = 50
        col = rng.integers(0, n_src, size=int(rowptr[-1])).astype(np.uint32)
        x = rand_u64(rng, n_src, 16)
        csr = cgb.csr_create(to_dev(rowptr), to_dev(col) if col.size else torch.empty(0, dtype=torch.int32, device="cuda"),
                             n_src)
        got = cgb.gather_sum(csr, to_dev(x))
        assert np.array_equal(to_np(got), oracle.gather_sum_csr(rowptr, col, x))
        csr.destroy()


def test_cora_small_worked_example(cgb, oracle):
    """party 0 of the cora_small shape; U_loc = [X2, X0], mirror block = [X0+X2, X2]."""
    from oracle import graph_index as gi

    edges = [(0, 1), (1, 0), (1, 2), (2, 1), (2, 3), (3, 2), (0, 2), (2, 0)]
    tiles, ivs = gi.build_all(edges, [0, 1, 0, 1], 2)
    iv = ivs[0]
    rng = np.random.default_rng(3)
    X = rand_u64(rng, 2, 3)  # rows: vertex 0, vertex 2
    rp, col = gi.csr_from_pos(iv["updateSrcVertexPos"][0], iv["updateDstVertexPos"][0], iv["localVertexPos"], iv["localVertexPos"])
    csr = cgb.csr_create(to_dev(rp, "cpu"), to_dev(col, "cpu"), 2)
    got = to_np(cgb.gather_sum(csr, to_dev(X)))
    assert np.array_equal(got, np.stack([X[1], X[0]]))
    rp, col = gi.csr_from_pos(iv["updateSrcVertexPos"][1], iv["updateDstVertexPos"][1], iv["localVertexPos"], ivs[1]["localVertexPos"])
    csr1 = cgb.csr_create(to_dev(rp, "cpu"), to_dev(col, "cpu"), 2)
    got = to_np(cgb.gather_sum(csr1, to_dev(X)))
    assert np.array_equal(got, np.stack([X[0] + X[1], X[1]]))
    csr.destroy(); csr1.destroy()


@pytest.mark.parametrize("D", [3, 16, 40, 130])
def test_expand_and_segsum_match_oracle(cgb, oracle, D):
    rng = np.random.default_rng(200 + D)
    n_dst, n_src = 800, 700
    rowptr, col = power_law_csr(rng, n_dst, n_src, 9000)
    x = rand_u64(rng, n_src, D)
    delta = rand_u64(rng, col.size, D)
    idx = col.copy()
    idx[::17] = 0xFFFFFFFF  # allowMissing positions
    for dl in (None, delta):
        want = oracle.expand_rows(idx, x, dl)
        got = cgb.expand_rows(to_dev(idx), to_dev(x), None if dl is None else to_dev(dl))
        assert np.array_equal(to_np(got), want)
    exp = oracle.expand_rows(col, x)
    for dup in (False, True):
        want = oracle.segsum(rowptr, exp, dup)
        got = cgb.segsum(to_dev(rowptr), to_dev(exp), dup)
        assert np.array_equal(to_np(got), want)
    # four-step dataflow of the reference == fused gather
    fused = cgb.gather_sum(cgb.csr_create(to_dev(rowptr), to_dev(col), n_src), to_dev(x))
    assert np.array_equal(to_np(fused), oracle.segsum(rowptr, exp, False))


def test_om_online_protocol_reconstructs(cgb, oracle):
    """Client y0 = A(x0 + (x1 - r)) + (A r - s), server y1 = s  =>  y0 + y1 = A x  (ssk.h:751-821 composite)."""
    rng = np.random.default_rng(9)
    n_dst, n_src, D = 500, 400, 16
    rowptr, col = power_law_csr(rng, n_dst, n_src, 6000)
    key = [45, 1, 2, 3, 4, 5, 6, 7]
    x = rand_u64(rng, n_src, D)
    x1 = rand_u64(rng, n_src, D)
    x0 = x - x1
    csr = cgb.csr_create(to_dev(rowptr), to_dev(col), n_src)
    r = cgb.prg_fill(key, 1, 0, n_src * D).view(n_src, D)
    s = cgb.prg_fill(key, 2, 0, n_dst * D).view(n_dst, D)
    delta = cgb.sub(cgb.gather_sum(csr, r), s)           # offline: A r - s
    m = cgb.prg_mask_sub(key, 1, 0, to_dev(x1))           # server message x1 - r
    y0 = cgb.gather_sum(csr, cgb.add(to_dev(x0), m), delta)
    y = to_np(y0) + to_np(s)
    assert np.array_equal(y, oracle.gather_sum_csr(rowptr, col, x))


def test_large_graph_properties(cgb):
    """At bench scale the oracle is too slow: check linearity and the checksum identity sum(y) = sum_e x[col[e]]."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(5)
    n, E, D = 200_000, 3_000_000, 16
    dst = torch.sort((torch.rand(E, device="cuda", generator=g) ** 3 * n).long().clamp_(max=n - 1)).values
    col = (torch.rand(E, device="cuda", generator=g) ** 2 * n).long().clamp_(max=n - 1).int()
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n), 0)
    csr = cgb.csr_create(rowptr.int(), col, n)
    a = torch.randint(-2**63, 2**63 - 1, (n, D), device="cuda", dtype=torch.int64, generator=g)
    b = torch.randint(-2**63, 2**63 - 1, (n, D), device="cuda", dtype=torch.int64, generator=g)
    ya, yb, yab = cgb.gather_sum(csr, a), cgb.gather_sum(csr, b), cgb.gather_sum(csr, a + b)
    assert torch.equal(ya + yb, yab)  # int64 wraps like Z_2^64
    assert torch.equal(ya.sum(0), a[col.long()].sum(0))
    csr.destroy()


@pytest.mark.parametrize("D", [7, 16, 64])
def test_gather_sum_blocks_matches_oracle(cgb, oracle, D):
    """Block outputs (one buffer per destination party; in deployment the buffers are peer memory over NVLink)."""
    import torch

    rng = np.random.default_rng(300 + D)
    n_dst, n_src = 2100, 1500
    rowptr, col = power_law_csr(rng, n_dst, n_src, 30000)
    x, delta = rand_u64(rng, n_src, D), rand_u64(rng, n_dst, D)
    csr = cgb.csr_create(to_dev(rowptr), to_dev(col), n_src)
    offsets = [0, 700, 700, 1500, 2100]  # includes an empty block
    bufs = [torch.empty((offsets[t + 1] - offsets[t], D), dtype=torch.int64, device="cuda") for t in range(4)]
    # empty tensors have a null data_ptr on some torch versions: point them at a dummy allocation
    dummy = torch.empty(2, dtype=torch.int64, device="cuda")
    ptrs = [b.data_ptr() if b.numel() else dummy.data_ptr() for b in bufs]
    for dl in (None, delta):
        want = oracle.gather_sum_csr(rowptr, col, x, dl)
        cgb.gather_sum_blocks(csr, to_dev(x), ptrs, offsets, None if dl is None else to_dev(dl))
        for t in range(4):
            assert np.array_equal(to_np(bufs[t]).reshape(-1, D), want[offsets[t]:offsets[t + 1]])
    csr.destroy()


def test_host_entry_points_sync_and_pipelined(cgb, oracle):
    """cgb_host_gather_sum and the pipelined cgb_host_gather_sum_async + cgb_host_sync with a different input per step
    (a slot-reuse or ordering bug would mix the steps)."""
    import ctypes as C

    import torch

    rng = np.random.default_rng(17)
    n_dst, n_src, D = 900, 800, 16
    rowptr, col = power_law_csr(rng, n_dst, n_src, 12000)
    csr = cgb.csr_create(to_dev(rowptr), to_dev(col), n_src)
    steps = 7
    xs = [rand_u64(rng, n_src, D) for _ in range(steps)]
    delta = rand_u64(rng, n_dst, D)
    hx = [torch.from_numpy(x.view(np.int64)).pin_memory() for x in xs]
    hy = [torch.zeros((n_dst, D), dtype=torch.int64).pin_memory() for _ in range(steps)]
    hd = torch.from_numpy(delta.view(np.int64)).pin_memory()
    lib = cgb.lib
    for i in range(steps):
        dl = C.c_void_p(hd.data_ptr()) if i % 2 else None
        cgb.check(lib.cgb_host_gather_sum_async(cgb.handle, csr.handle, C.c_void_p(hx[i].data_ptr()), dl,
                                                C.c_void_p(hy[i].data_ptr()), D))
    cgb.check(lib.cgb_host_sync(cgb.handle))
    for i in range(steps):
        want = oracle.gather_sum_csr(rowptr, col, xs[i], delta if i % 2 else None)
        assert np.array_equal(hy[i].numpy().view(np.uint64), want), i
    out = torch.zeros((n_dst, D), dtype=torch.int64).pin_memory()
    cgb.check(lib.cgb_host_gather_sum(cgb.handle, csr.handle, C.c_void_p(hx[0].data_ptr()), None, C.c_void_p(out.data_ptr()), D))
    assert np.array_equal(out.numpy().view(np.uint64), oracle.gather_sum_csr(rowptr, col, xs[0]))
    csr.destroy()


@pytest.mark.parametrize("D", [2, 16, 33])
def test_gather_sum_chunk_boundaries(cgb, oracle, D):
    """Rows that start / end exactly on the 64-edge chunk boundaries of the edge-balanced schedule, rows spanning several
    chunks, runs of empty rows between them, an empty first and last row, and the whole thing repeated with delta."""
    rng = np.random.default_rng(D)
    degs = [0, 64, 64, 1, 63, 65, 0, 0, 128, 1, 1, 62, 200, 0, 64, 3, 61, 129, 0, 500, 64, 0]
    rowptr = np.zeros(len(degs) + 1, dtype=np.uint32)
    rowptr[1:] = np.cumsum(degs)
    n_src = 97
    col = rng.integers(0, n_src, size=int(rowptr[-1])).astype(np.uint32)
    x, delta = rand_u64(rng, n_src, D), rand_u64(rng, len(degs), D)
    csr = cgb.csr_create(to_dev(rowptr, "cpu"), to_dev(col, "cpu"), n_src)
    for dl in (None, delta):
        want = oracle.gather_sum_csr(rowptr, col, x, dl)
        for _ in range(3):  # arrival counters must reset themselves between launches
            got = cgb.gather_sum(csr, to_dev(x), None if dl is None else to_dev(dl))
            assert np.array_equal(to_np(got), want)
    csr.destroy()


def test_gather_sum_randomised_shapes(cgb, oracle):
    rng = np.random.default_rng(2024)
    for trial in range(25):
        n_dst, n_src = int(rng.integers(1, 400)), int(rng.integers(1, 300))
        D = int(rng.integers(1, 90))
        deg = (rng.pareto(0.8, size=n_dst) * rng.integers(0, 6)).astype(np.int64)
        deg[rng.random(n_dst) < 0.3] = 0
        rowptr = np.zeros(n_dst + 1, dtype=np.uint32)
        rowptr[1:] = np.cumsum(np.minimum(deg, 3000))
        col = rng.integers(0, n_src, size=int(rowptr[-1])).astype(np.uint32)
        x = rand_u64(rng, n_src, D)
        delta = rand_u64(rng, n_dst, D) if trial % 2 else None
        import torch

        dcol = to_dev(col) if col.size else torch.empty(0, dtype=torch.int32, device="cuda")
        csr = cgb.csr_create(to_dev(rowptr), dcol, n_src)
        got = cgb.gather_sum(csr, to_dev(x), None if delta is None else to_dev(delta))
        assert np.array_equal(to_np(got), oracle.gather_sum_csr(rowptr, col, x, delta)), (trial, n_dst, n_src, D)
        csr.destroy()


@pytest.mark.parametrize("n_words,n_ctas", [(2, 0), (254, 1), (2 * 128 * 8 * 3 + 6, 3), (1 << 20, 0), (1_000_002, 64)])
def test_peer_copy_is_a_copy(cgb, n_words, n_ctas):
    """cgb_peer_copy (the push / pull transport of the mirror-update exchange) on local buffers: every word arrives, the
    words beyond the range are untouched, sizes that are not a multiple of the unrolled tile take the tail loop."""
    import torch

    rng = np.random.default_rng(n_words)
    src = to_dev(rand_u64(rng, n_words))
    dst = torch.full((n_words + 4,), -1, dtype=torch.int64, device="cuda")
    cgb.peer_copy(dst.data_ptr(), src.data_ptr(), n_words * 8, n_ctas)
    torch.cuda.synchronize()
    assert torch.equal(dst[:n_words], src)
    assert bool((dst[n_words:] == -1).all())
    with pytest.raises(Exception):
        cgb.peer_copy(dst.data_ptr() + 8, src.data_ptr(), 16, 0)  # 16-byte alignment is part of the contract


@pytest.mark.parametrize("env", [{"CGB_GATHER_VEC": "2"}, {"CGB_GATHER_VEC": "4", "CGB_GATHER_IPL": "1"},
                                 {"CGB_GATHER_IMPL": "rows"}, {"CGB_GATHER_IMPL": "async", "CGB_GATHER_VEC": "2"},
                                 {"CGB_CHUNK_SHIFT": "7"}, {"CGB_CHUNK_SHIFT": "5", "CGB_GATHER_VEC": "2"}])
def test_gather_alternative_kernels_in_subprocess(env):
    """The default is the 256-bit edge-balanced kernel with 64-edge chunks (128 from 16M edges on); the 128-bit, row-schedule
    and cp.async variants and other chunk sizes stay selectable for A/B measurements (environment read once per process), so
    they are checked against the oracle in a child process."""
    import os
    import subprocess
    import sys

    code = r"""
import numpy as np, sys
sys.path.insert(0, %r)
import cognn_b200
from oracle import pyoracle
from tests.util import power_law_csr, rand_u64, to_dev, to_np
pyoracle.build()
ctx = cognn_b200.Context(0)
for D in (4, 16, 40, 64, 128):
    rng = np.random.default_rng(D)
    rowptr, col = power_law_csr(rng, 2000, 1500, 30000)
    x, delta = rand_u64(rng, 1500, D), rand_u64(rng, 2000, D)
    csr = ctx.csr_create(to_dev(rowptr, "cpu"), to_dev(col, "cpu"), 1500)
    for dl in (None, delta):
        got = to_np(ctx.gather_sum(csr, to_dev(x), None if dl is None else to_dev(dl)))
        assert np.array_equal(got, pyoracle.gather_sum_csr(rowptr, col, x, dl)), D
print("ok")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("D", [1, 2, 7, 16, 40, 64])
def test_gather_sum_compact_and_scatter_add(cgb, oracle, D):
    """The compact mirror-update block (one row per destination that has an edge) and its consumer: scattering / adding it
    into the vertex rows reproduces the dense gather, with and without the offline correlation delta."""
    import torch

    rng = np.random.default_rng(400 + D)
    n_dst, n_src = 2300, 1700
    rowptr, col = power_law_csr(rng, n_dst, n_src, 30000)
    deg = np.diff(rowptr.astype(np.int64))
    nz = np.nonzero(deg)[0]
    assert 0 < nz.size < n_dst
    x = rand_u64(rng, n_src, D)
    csr = cgb.csr_create(to_dev(rowptr), to_dev(col), n_src)
    assert csr.n_nonempty == nz.size
    assert np.array_equal(to_np(csr.nonempty_rows()).astype(np.int64), nz)
    dense = oracle.gather_sum_csr(rowptr, col, x)
    delta_c = rand_u64(rng, nz.size, D)
    for dl in (None, delta_c):
        got = cgb.gather_sum_compact(csr, to_dev(x), None if dl is None else to_dev(dl))
        want = dense[nz] + (0 if dl is None else dl)
        for _ in range(2):  # second launch: arrival counters reset themselves
            assert np.array_equal(to_np(got), want)
            got = cgb.gather_sum_compact(csr, to_dev(x), None if dl is None else to_dev(dl))
    # consumer side: v += block (GatherComp addition), and the assign form
    block = cgb.gather_sum_compact(csr, to_dev(x))
    base = rand_u64(rng, n_dst, D)
    v = to_dev(base)
    cgb.scatter_add_rows(csr.nonempty_rows(), block, v)
    assert np.array_equal(to_np(v), base + dense)
    v = to_dev(base)
    cgb.scatter_add_rows(csr.nonempty_rows(), block, v, assign=True)
    want = base.copy()
    want[nz] = dense[nz]
    assert np.array_equal(to_np(v), want)
    # raw-address source (what a peer's staging buffer looks like) and a tiny grid
    v = to_dev(base)
    cgb.scatter_add_rows(csr.nonempty_rows(), int(block.data_ptr()), v, n=nz.size, n_ctas=3)
    assert np.array_equal(to_np(v), base + dense)
    torch.cuda.synchronize()
    csr.destroy()


def test_gather_sum_compact_edge_cases(cgb, oracle):
    import torch

    rng = np.random.default_rng(11)
    # no edges at all; one row; rows cut by chunk boundaries with empty rows in between
    for degs in ([0, 0, 0], [5], [0, 64, 0, 65, 200, 0, 0, 1, 63, 500, 0]):
        rowptr = np.zeros(len(degs) + 1, dtype=np.uint32)
        rowptr[1:] = np.cumsum(degs)
        col = rng.integers(0, 40, size=int(rowptr[-1])).astype(np.uint32)
        x = rand_u64(rng, 40, 16)
        dcol = to_dev(col) if col.size else torch.empty(0, dtype=torch.int32, device="cuda")
        csr = cgb.csr_create(to_dev(rowptr), dcol, 40)
        nz = np.nonzero(np.array(degs))[0]
        assert csr.n_nonempty == nz.size
        got = cgb.gather_sum_compact(csr, to_dev(x))
        assert got.shape[0] == nz.size
        if nz.size:
            assert np.array_equal(to_np(got), oracle.gather_sum_csr(rowptr, col, x)[nz])
        csr.destroy()


@pytest.mark.parametrize("mode", [0, 1])
def test_flags_signal_then_wait(cgb, mode):
    """cgb_flag_signal / cgb_flag_wait on one stream: a wait for a value already signalled (or exceeded) passes; the polling
    form gives up on a flag nobody raises and reports it instead of hanging."""
    import time

    import torch

    flag = torch.zeros(4, dtype=torch.int32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    cgb.flag_signal(flag.data_ptr(), 3)
    cgb.flag_wait(flag.data_ptr(), 3, mode, err.data_ptr())
    cgb.flag_wait(flag.data_ptr(), 2, mode, err.data_ptr())  # flags only grow: an older value is satisfied too
    marker = torch.ones(1, device="cuda")  # enqueued behind the waits on the same (current) stream
    torch.cuda.synchronize()
    assert int(flag[0]) == 3 and int(err[0]) == 0 and float(marker[0]) == 1.0
    if mode == 1:
        t0 = time.time()
        cgb.flag_wait(flag.data_ptr() + 4, 1, 1, err.data_ptr())
        torch.cuda.synchronize()
        assert int(err[0]) == 1 and time.time() - t0 < 30.0


@pytest.mark.parametrize("D", [2, 16, 40, 64])
def test_gather_sum_signal_blocks_and_flags(cgb, oracle, D):
    """The fused gather + signalling launch: rows cut into dense and compact blocks (an empty block and an all-empty-rows block
    included), every block in its own buffer, every flag raised exactly with the launch's value; repeated launches reuse the
    self-resetting completion counters."""
    import torch

    rng = np.random.default_rng(500 + D)
    n_dst, n_src = 2600, 1900
    rowptr, col = power_law_csr(rng, n_dst, n_src, 36000)
    deg = np.diff(rowptr.astype(np.int64)).copy()
    # make rows [900, 1000) empty so that one block has no edge at all
    keep = np.ones(n_dst, dtype=bool)
    keep[900:1000] = False
    e_keep = np.repeat(keep, deg)
    col = col[e_keep]
    deg[~keep] = 0
    rowptr = np.zeros(n_dst + 1, dtype=np.uint32)
    rowptr[1:] = np.cumsum(deg)
    x = rand_u64(rng, n_src, D)
    dense = oracle.gather_sum_csr(rowptr, col, x)
    csr = cgb.csr_create(to_dev(rowptr), to_dev(col), n_src)
    offsets = [0, 700, 700, 900, 1000, 1777, 2600]  # empty block, block of empty rows, ragged cuts
    compact = [False, True, True, True, False, True]
    nb = len(compact)
    nz_in = [np.nonzero(deg[offsets[b]:offsets[b + 1]])[0] + offsets[b] for b in range(nb)]
    bufs = [torch.full((max(1, (nz_in[b].size if compact[b] else offsets[b + 1] - offsets[b])), D), -1, dtype=torch.int64, device="cuda")
            for b in range(nb)]
    flags = torch.zeros(nb, dtype=torch.int32, device="cuda")
    for value in (1, 2, 7):
        cgb.gather_sum_signal(csr, to_dev(x), offsets, [t.data_ptr() for t in bufs], compact,
                              [flags.data_ptr() + 4 * b for b in range(nb)], value)
        torch.cuda.synchronize()
        assert flags.tolist() == [value] * nb
        for b in range(nb):
            got = to_np(bufs[b])
            if compact[b]:
                assert np.array_equal(got[:nz_in[b].size], dense[nz_in[b]]), (b, value)
            else:
                assert np.array_equal(got[:offsets[b + 1] - offsets[b]], dense[offsets[b]:offsets[b + 1]]), (b, value)
    # no flags requested: plain blocked output
    for t in bufs:
        t.fill_(-1)
    cgb.gather_sum_signal(csr, to_dev(x), offsets, [t.data_ptr() for t in bufs], compact, [0] * nb, 9)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(bufs[0])[:700], dense[:700]) and flags.tolist() == [7] * nb
    csr.destroy()


@pytest.mark.parametrize("n_edges,D", [(300_000, 16), (600_000, 7), (1_500_000, 40), (2_500_000, 16)])
def test_gather_sum_every_chunk_size_class(cgb, oracle, n_edges, D):
    """The chunk size follows the edge count (16 / 32 / 64 / 128 edges below 2^18 / 2^21 / 2^24 / above): the small tests above
    run 16-edge chunks, the 2^24- and 10^8-edge tests 128; these sizes run the 32- and 64-edge classes, with hub rows that span
    hundreds of chunks, empty rows, and delta, against the oracle on every row."""
    rng = np.random.default_rng(n_edges % 1000 + D)
    n_dst, n_src = n_edges // 12, n_edges // 10
    rowptr, col = power_law_csr(rng, n_dst, n_src, n_edges)
    deg = np.diff(rowptr.astype(np.int64))
    assert deg.max() > 4000 and (deg == 0).any()
    x, delta = rand_u64(rng, n_src, D), rand_u64(rng, n_dst, D)
    csr = cgb.csr_create(to_dev(rowptr, "cpu"), to_dev(col, "cpu"), n_src)
    dx, dd = to_dev(x), to_dev(delta)
    for use_delta in (False, True):
        want = oracle.gather_sum_csr(rowptr, col, x, delta if use_delta else None)
        for _ in range(2):  # the second launch reuses the self-resetting arrival counters
            got = cgb.gather_sum(csr, dx, dd if use_delta else None)
            assert np.array_equal(to_np(got), want)
    # the compact form on the same CSR
    nz = to_np(csr.nonempty_rows()).astype(np.int64)
    want = oracle.gather_sum_csr(rowptr, col, x, None)
    got = cgb.gather_sum_compact(csr, dx)
    assert np.array_equal(to_np(got), want[nz])
    csr.destroy()
