"""GPU parity (bit exact) of ops (2), (3), (4) against the CPU oracle, through the C ABI."""
import json
import os

import numpy as np
import pytest

from tests.util import rand_u64, to_dev, to_np

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("M,K,N", [(1, 1, 1), (5, 3, 2), (130, 17, 7), (257, 1433, 16), (300, 64, 40), (129, 500, 3),
                                    (64, 128, 256), (1000, 33, 65), (16, 21168, 40), (7, 1354, 16), (16, 1354, 7),
                                    (32, 5000, 64), (3, 700, 2)])
def test_matmul_matches_oracle(cgb, oracle, M, K, N):
    rng = np.random.default_rng(M * 7 + K * 3 + N)
    A, B = rand_u64(rng, M, K), rand_u64(rng, K, N)
    want = oracle.matmul(A, B)
    assert np.array_equal(to_np(cgb.matmul(to_dev(A), to_dev(B))), want)
    At = np.ascontiguousarray(A.T)
    assert np.array_equal(to_np(cgb.matmul(to_dev(At), to_dev(B), transA=True)), want)
    C0 = rand_u64(rng, M, N)
    out = to_dev(C0)
    cgb.matmul(to_dev(A), to_dev(B), out=out, accumulate=True)
    assert np.array_equal(to_np(out), oracle.matmul(A, B, C_in=C0))


def test_matmul_split_k_weight_gradient_shape(cgb, oracle):
    # d = h_t * v (gcn.h:671): K = N_p is long, the output F x H tile is small -> split-K with 64-bit atomics
    rng = np.random.default_rng(11)
    n_p, F, H = 5000, 128, 16
    X, G = rand_u64(rng, n_p, F), rand_u64(rng, n_p, H)
    want = oracle.matmul(X, G, transA=True)
    assert np.array_equal(to_np(cgb.matmul(to_dev(X), to_dev(G), transA=True)), want)
    C0 = rand_u64(rng, F, H)
    out = to_dev(C0)
    cgb.matmul(to_dev(X), to_dev(G), transA=True, out=out, accumulate=True)
    assert np.array_equal(to_np(out), oracle.matmul(X, G, transA=True, C_in=C0))


@pytest.mark.parametrize("M,K,N", [(9, 6, 4), (700, 500, 16), (64, 4000, 8), (300, 128, 256)])
@pytest.mark.parametrize("f", [-1, 16])
def test_beaver_matmul_finish_matches_oracle(cgb, oracle, M, K, N, f):
    rng = np.random.default_rng(M + K + N)
    E, F, U, V, Z = (rand_u64(rng, M, K), rand_u64(rng, K, N), rand_u64(rng, M, K), rand_u64(rng, K, N),
                     rand_u64(rng, M, N))
    for share in (0, 1):
        want = oracle.beaver_matmul_finish(E, F, U, V, Z, share, f)
        got = cgb.beaver_matmul_finish(to_dev(E), to_dev(F), to_dev(U), to_dev(V), to_dev(Z), share, f)
        assert np.array_equal(to_np(got), want)


def test_elementwise_match_oracle(cgb, oracle):
    rng = np.random.default_rng(21)
    for n in (1, 2, 3, 1000, 4097):
        a, b = rand_u64(rng, n), rand_u64(rng, n)
        da, db = to_dev(a), to_dev(b)
        assert np.array_equal(to_np(cgb.add(da, db)), oracle.add(a, b))
        assert np.array_equal(to_np(cgb.sub(da, db)), oracle.sub(a, b))
        for share in (0, 1):
            assert np.array_equal(to_np(cgb.trunc(da, share, 16)), oracle.trunc(a, 16, share))
            assert np.array_equal(to_np(cgb.scale_public(da, 12345, share, 16)), oracle.scale_public(a, 12345, 16, share))
            assert np.array_equal(to_np(cgb.apply_gradient(da, db, 32768, share, 16)),
                                  oracle.apply_gradient(a, b, 32768, 16, share))
        # unaligned views (odd element offset -> scalar path) and in-place (reference aliases in/out, gcn.h:676)
        if n > 3:
            assert np.array_equal(to_np(cgb.add(da[1:], db[1:])), oracle.add(a[1:], b[1:]))
            tmp = da.clone()
            cgb.scale_public(tmp, 777, 0, 16, out=tmp)
            assert np.array_equal(to_np(tmp), oracle.scale_public(a, 777, 16, 0))


@pytest.mark.parametrize("rows,D", [(1, 1), (11, 6), (1354, 16), (500, 7)])
def test_rowmul_cond_transpose_match_oracle(cgb, oracle, rows, D):
    rng = np.random.default_rng(rows + D)
    e, a, c = rand_u64(rng, rows, D), rand_u64(rng, rows, D), rand_u64(rng, rows, D)
    fv, b = rand_u64(rng, rows), rand_u64(rng, rows)
    for share in (0, 1):
        for f in (-1, 16):
            want = oracle.rowmul_beaver_finish(e, fv, a, b, c, share, f)
            got = cgb.rowmul_beaver_finish(to_dev(e), to_dev(fv), to_dev(a), to_dev(b), to_dev(c), share, f)
            assert np.array_equal(to_np(got), want)
    cond = (rng.integers(0, 2, size=rows)).astype(np.uint8)
    assert np.array_equal(to_np(cgb.cond_add(to_dev(e), to_dev(a), to_dev(cond))), oracle.cond_add(e, a, cond))
    assert np.array_equal(to_np(cgb.transpose(to_dev(e))), oracle.transpose(e))


def test_encode_split_open_match_oracle(cgb, oracle):
    rng = np.random.default_rng(33)
    x = rng.normal(size=(123, 17)) * 50
    x.ravel()[:4] = [0.0, -0.0, 1.0 / 3.0, -1.0 / 3.0]
    key = [9, 8, 7, 6, 5, 4, 3, 2]
    assert np.array_equal(to_np(cgb.encode(to_dev(x))), oracle.encode(x, 16))
    s0, s1 = cgb.share_split(to_dev(x), key, stream=77, word_offset=5)
    w0, w1 = oracle.share_split(x, 16, key, 77, 5)
    assert np.array_equal(to_np(s0), w0) and np.array_equal(to_np(s1), w1)
    assert np.array_equal(to_np(cgb.open_decode(s0, s1)), oracle.open_decode(w0, w1, 16))
    assert np.array_equal(to_np(cgb.decode(s0)), oracle.decode(w0, 16))


def test_prg_matches_openssl_golden_and_oracle(cgb, oracle):
    cases = json.load(open(os.path.join(HERE, "golden", "chacha20_openssl.json")))["cases"]
    for c in cases:
        words = np.array([int(w) for w in c["words"]], dtype=np.uint64)
        got = cgb.prg_fill(c["key"], c["stream"], c["word_offset"], words.size)
        assert np.array_equal(to_np(got), words)
        got = cgb.prg_fill(c["key"], c["stream"], c["word_offset"] + 3, words.size - 5)
        assert np.array_equal(to_np(got), words[3:-2])
    key = [45, 0, 1, 2, 3, 4, 5, 6]
    for off, n in [(0, 1), (7, 1), (8, 300), (13, 100_003)]:
        want = oracle.prg_fill(key, 0x1234567890, off, n)
        assert np.array_equal(to_np(cgb.prg_fill(key, 0x1234567890, off, n)), want)
        x = rand_u64(np.random.default_rng(n), n)
        assert np.array_equal(to_np(cgb.prg_mask_sub(key, 0x1234567890, off, to_dev(x))), x - want)
    assert cgb.prg_fill(key, 1, 0, 0).numel() == 0


def test_ideal_functionality_standins(cgb):
    """2PC-residual stand-ins (not secure): exact integer ReLU / ReLU' on reconstructed values, as oracle/epoch.py."""
    rng = np.random.default_rng(5)
    n = 5001
    v = (rng.normal(size=n) * 1000).astype(np.int64).astype(np.uint64)
    z = (rng.normal(size=n) * 1000).astype(np.int64).astype(np.uint64)
    v[:3] = [0, 1, np.uint64(2**64 - 1)]
    a1, z1 = rand_u64(rng, n), rand_u64(rng, n)
    a0, z0 = v - a1, z - z1
    got = to_np(cgb.ideal_relu(to_dev(a0), to_dev(a1)))
    assert np.array_equal(got, np.where(v.astype(np.int64) > 0, v, np.uint64(0)))
    got = to_np(cgb.ideal_relu_grad(to_dev(a0), to_dev(a1), to_dev(z0), to_dev(z1)))
    assert np.array_equal(got, np.where(z.astype(np.int64) > 0, v, np.uint64(0)))


def test_prg_stream_bias_is_read_on_the_device(cgb, oracle):
    """cgb_ctx_set_prg_stream_bias: stream id = argument + device word, read when the kernel runs (graph replays)."""
    import torch

    key = list(range(11, 19))
    bias = torch.zeros(1, dtype=torch.int64, device="cuda")
    x = rand_u64(np.random.default_rng(5), 1000)
    cgb.set_prg_stream_bias(bias)
    try:
        for b in (0, 6 << 16, (12 << 16) + 3):
            bias.fill_(b)
            assert np.array_equal(to_np(cgb.prg_fill(key, 77 << 48, 9, 1000)), oracle.prg_fill(key, (77 << 48) + b, 9, 1000))
            assert np.array_equal(to_np(cgb.prg_mask_sub(key, 5, 0, to_dev(x))), x - oracle.prg_fill(key, 5 + b, 0, 1000))
    finally:
        cgb.set_prg_stream_bias(None)
    assert np.array_equal(to_np(cgb.prg_fill(key, 5, 0, 64)), oracle.prg_fill(key, 5, 0, 64))


@pytest.mark.parametrize("n,C,spread", [(1, 3, 1.0), (257, 7, 3.0), (5000, 40, 8.0), (300, 70, 40.0)])
def test_ideal_softmax_bit_exact(cgb, oracle, n, C, spread):
    """Device stand-in of the prediction layer == CPU oracle, bit for bit (exp is restated from IEEE + - * on both sides).
    Logits up to +-40*4 fixed-point units apart exercise the exp underflow-to-zero range; ties and equal rows too."""
    import torch

    rng = np.random.default_rng(n * 100 + C)
    f = 16
    z = (rng.normal(0, spread, size=(n, C)) * (1 << f)).astype(np.int64)
    z[0] = z[0, 0]            # a constant row: uniform probabilities
    if n > 2:
        z[1, 0] = 700 << f    # one dominant class: the others underflow
        z[2] = -(1 << 40)     # large negative, all equal
    z = z.view(np.uint64)
    z1 = rand_u64(rng, n, C)
    z0 = z - z1
    labels = rng.integers(0, C, size=n).astype(np.int32)
    train = n * 2 // 5
    want_P, want_d = oracle.ideal_softmax(z0, z1, labels, train, f)
    P, d = cgb.ideal_softmax(to_dev(z0), to_dev(z1), torch.from_numpy(labels).cuda(), train, f)
    assert np.array_equal(to_np(P), want_P)
    assert np.array_equal(to_np(d), want_d)
    assert not to_np(d)[train:].any()
    # probabilities sum to one up to the C truncations
    tot = to_np(P).astype(np.int64).sum(axis=1)
    assert ((1 << f) - C <= tot).all() and (tot <= (1 << f)).all()


def test_ideal_softmax_matches_committed_golden(cgb):
    """The CUDA stand-in against tests/golden/ideal_softmax.json directly (no oracle run in between)."""
    import json
    import os

    import torch

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ideal_softmax.json")))
    shape = tuple(g["shape"])
    z0 = np.array(g["z0"], dtype=np.uint64).reshape(shape)
    z1 = np.array(g["z1"], dtype=np.uint64).reshape(shape)
    P, d = cgb.ideal_softmax(to_dev(z0), to_dev(z1), torch.tensor(g["labels"], dtype=torch.int32, device="cuda"),
                             g["train_rows"], g["f"])
    assert to_np(P).ravel().tolist() == g["P"] and to_np(d).ravel().tolist() == g["pmy"]


def test_sum_n_matches_oracle(cgb, oracle):
    rng = np.random.default_rng(8)
    for n in (1, 7, 4096, 100_001):
        xs = [rand_u64(rng, n) for _ in range(5)]
        want = xs[0].copy()
        for x in xs[1:]:
            want = oracle.add(want, x)
        devs = [to_dev(x) for x in xs]
        assert np.array_equal(to_np(cgb.sum_n(devs)), want)
        assert np.array_equal(to_np(cgb.sum_n(devs[:1])), xs[0])
        cgb.sum_n(devs, out=devs[0])  # in place on the first input
        assert np.array_equal(to_np(devs[0]), want)


def test_matmul_properties_at_sweep_size(cgb):
    """configs[4] size (2^18 x 512 x 256), too slow for the oracle: linearity in A and the transposed-storage identity."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(3)
    M, K, N = 1 << 18, 512, 256
    A = torch.randint(-2**63, 2**63 - 1, (M, K), device="cuda", dtype=torch.int64, generator=g)
    B = torch.randint(-2**63, 2**63 - 1, (M, K), device="cuda", dtype=torch.int64, generator=g)
    W = torch.randint(-2**63, 2**63 - 1, (K, N), device="cuda", dtype=torch.int64, generator=g)
    ya, yb, yab = cgb.matmul(A, W), cgb.matmul(B, W), cgb.matmul(A + B, W)
    assert torch.equal(ya + yb, yab)
    # a few rows against exact integer arithmetic on the host
    rows = [0, 12345, M - 1]
    Wl = [[int(v) for v in r] for r in W.cpu().tolist()]
    for r in rows:
        a = [int(v) for v in A[r].cpu().tolist()]
        want = [sum(a[k] * Wl[k][j] for k in range(K)) % (1 << 64) for j in (0, 7, N - 1)]
        got = [int(ya[r, j].item()) % (1 << 64) for j in (0, 7, N - 1)]
        assert got == want
    # weight-gradient form: X^T G with X stored n x F (transA) equals the explicit transpose
    X = A[: 1 << 16, :128].contiguous()
    G = B[: 1 << 16, :16].contiguous()
    assert torch.equal(cgb.matmul(X, G, transA=True), cgb.matmul(cgb.transpose(X), G))


def test_prg_stream_is_position_addressable_at_scale(cgb):
    key = [1, 2, 3, 4, 5, 6, 7, 8]
    n = 40_000_003
    whole = cgb.prg_fill(key, 99, 5, n)
    cut = 17_000_001
    a, b = cgb.prg_fill(key, 99, 5, cut), cgb.prg_fill(key, 99, 5 + cut, n - cut)
    import torch

    assert torch.equal(whole[:cut], a) and torch.equal(whole[cut:], b)
    assert not torch.equal(cgb.prg_fill(key, 100, 5, 1000), whole[:1000])


@pytest.mark.parametrize("M,K,N", [(128, 32, 64), (128, 64, 64), (256, 128, 128), (300, 100, 70), (1000, 512, 256),
                                    (129, 1433, 16), (2048, 4096, 64), (500, 5000, 40)])
def test_tensor_core_matmul_matches_oracle(cgb, oracle, M, K, N, monkeypatch):
    """tcgen05 kind::i8 limb path (CGB_MATMUL_IMPL=tc) bit-exact against the oracle, incl. padding in every dimension,
    K > 4096 (chunked accumulation), transposed storage, accumulate, and the fused Beaver finish."""
    monkeypatch.setenv("CGB_MATMUL_IMPL", "tc")
    rng = np.random.default_rng(M + 3 * K + 7 * N)
    A, B = rand_u64(rng, M, K), rand_u64(rng, K, N)
    # extreme limbs: all-ones words stress the diagonal accumulators
    A[0, :] = np.uint64(2**64 - 1)
    B[:, 0] = np.uint64(2**64 - 1)
    want = oracle.matmul(A, B)
    assert np.array_equal(to_np(cgb.matmul(to_dev(A), to_dev(B))), want)
    At = np.ascontiguousarray(A.T)
    assert np.array_equal(to_np(cgb.matmul(to_dev(At), to_dev(B), transA=True)), want)
    C0 = rand_u64(rng, M, N)
    out = to_dev(C0)
    cgb.matmul(to_dev(A), to_dev(B), out=out, accumulate=True)
    assert np.array_equal(to_np(out), oracle.matmul(A, B, C_in=C0))
    if K <= 2048:
        E, F, U, V, Z = A, B, rand_u64(rng, M, K), rand_u64(rng, K, N), rand_u64(rng, M, N)
        for share in (0, 1):
            for f in (-1, 16):
                want = oracle.beaver_matmul_finish(E, F, U, V, Z, share, f)
                got = cgb.beaver_matmul_finish(to_dev(E), to_dev(F), to_dev(U), to_dev(V), to_dev(Z), share, f)
                assert np.array_equal(to_np(got), want), (share, f)


# ---- fused forms the engine issues (one launch where the plain forms take two or three) ------------------------------------
@pytest.mark.parametrize("M,K,N", [(9, 6, 4), (700, 500, 16), (1433, 1354, 16), (300, 128, 256), (16, 9000, 40), (16, 1354, 7)])
@pytest.mark.parametrize("f", [-1, 16])
def test_beaver_matmul_finish_open_matches_oracle(cgb, oracle, M, K, N, f):
    rng = np.random.default_rng(M + K + N + 1)
    U, V, Z = rand_u64(rng, M, K), rand_u64(rng, K, N), rand_u64(rng, M, N)
    mine, peer = rand_u64(rng, M * K + K * N), rand_u64(rng, M * K + K * N)
    opened = mine + peer
    E, F = opened[:M * K].reshape(M, K), opened[M * K:].reshape(K, N)
    for share in (0, 1):
        want = oracle.beaver_matmul_finish(E, F, U, V, Z, share, f)
        m = to_dev(mine)
        got = cgb.beaver_matmul_finish_open(m, to_dev(peer), to_dev(U), to_dev(V), to_dev(Z), share, f)
        assert np.array_equal(to_np(got), want)
        assert np.array_equal(to_np(m), opened)  # the message is opened in place


@pytest.mark.parametrize("rows,D", [(1, 1), (11, 6), (1354, 16), (500, 7), (0, 4)])
def test_rowmul_finish_open_and_sub_pair_match_oracle(cgb, oracle, rows, D):
    rng = np.random.default_rng(rows * 31 + D)
    a, c, x = rand_u64(rng, rows, D), rand_u64(rng, rows, D), rand_u64(rng, rows, D)
    b, sc = rand_u64(rng, rows), rand_u64(rng, rows)
    mine, peer = rand_u64(rng, rows * D + rows), rand_u64(rng, rows * D + rows)
    opened = mine + peer
    e, fv = opened[:rows * D].reshape(rows, D), opened[rows * D:]
    if rows:
        for share in (0, 1):
            for f in (-1, 16):
                want = oracle.rowmul_beaver_finish(e, fv, a, b, c, share, f)
                got = cgb.rowmul_beaver_finish_open(to_dev(mine), to_dev(peer), to_dev(a), to_dev(b), to_dev(c), share, f)
                assert np.array_equal(to_np(got), want)
    # the message builder: [x - a | scaler - b], and the helper's form with a zero scaler share
    got = to_np(cgb.sub_pair(to_dev(x), to_dev(a), to_dev(sc), to_dev(b)))
    assert np.array_equal(got, np.concatenate([(x - a).ravel(), sc - b]))
    got = to_np(cgb.sub_pair(to_dev(x), to_dev(a), None, to_dev(b)))
    assert np.array_equal(got, np.concatenate([(x - a).ravel(), (0 - b.astype(np.uint64)).astype(np.uint64)]))


@pytest.mark.parametrize("n", [1, 7, 8, 300, 100_003])
@pytest.mark.parametrize("n_streams,n_in", [(1, 2), (3, 2), (1, 4), (0, 3), (2, 0), (7, 8)])
def test_prg_sum_matches_oracle(cgb, oracle, n, n_streams, n_in):
    key = [45, 0, 1, 2, 3, 4, 5, 6]
    rng = np.random.default_rng(n + 17 * n_streams + n_in)
    streams = [int(s) for s in rng.integers(1, 2**62, size=n_streams)]
    ins = [rand_u64(rng, n) for _ in range(n_in)]
    want = np.zeros(n, dtype=np.uint64)
    for s in streams:
        want = want + oracle.prg_fill(key, s, 0, n)
    for x in ins:
        want = want + x
    got = cgb.prg_sum(key, streams, [to_dev(x) for x in ins], n_words=n)
    assert np.array_equal(to_np(got), want)
    if n_in:  # out may alias an input
        d = [to_dev(x) for x in ins]
        cgb.prg_sum(key, streams, d, n_words=n, out=d[0])
        assert np.array_equal(to_np(d[0]), want)


def test_copy_segments(cgb):
    rng = np.random.default_rng(5)
    sizes = [1, 0, 7, 4096, 100_003] + [33] * 20  # > 16 segments: two launches; a zero-length one in the middle
    srcs = [to_dev(rand_u64(rng, n)) for n in sizes]
    dsts = [cgb.empty(n) for n in sizes]
    before = cgb.launches
    cgb.copy_segments(dsts, srcs)
    assert cgb.launches - before == 2
    for d, s in zip(dsts, srcs):
        assert np.array_equal(to_np(d), to_np(s))
    # unaligned views take the word-wise path
    a, b = to_dev(rand_u64(rng, 101)), cgb.empty(101)
    cgb.copy_segments([b[1:]], [a[1:]])
    assert np.array_equal(to_np(b[1:]), to_np(a[1:]))


def test_scale_apply_avg_and_relu_reshare_match_oracle(cgb, oracle):
    rng = np.random.default_rng(77)
    n = 1433 * 16 + 3
    W, d = rand_u64(rng, n), rand_u64(rng, n)
    gs, lr = 0x1F3, 0x28F
    for share in (0, 1):
        ds = oracle.scale_public(d, gs, 16, share)
        want_W = oracle.apply_gradient(W, ds, lr, 16, share)
        dW, dd = to_dev(W), to_dev(d)
        cgb.scale_apply_gradient(dW, dd, gs, lr, share)
        assert np.array_equal(to_np(dd), ds) and np.array_equal(to_np(dW), want_W)
        # average of 3 replicas' shares into 3 places, one of them an input
        ins = [rand_u64(rng, n) for _ in range(3)]
        want = oracle.scale_public(ins[0] + ins[1] + ins[2], 0x5555, 16, share)
        dins = [to_dev(x) for x in ins]
        outs = [cgb.empty(n), dins[0], dins[2]]
        cgb.avg_public(dins, 0x5555, outs, share)
        for o in outs:
            assert np.array_equal(to_np(o), want)
    # stand-in + re-share: relu(a0 + a1) - PRG, and the gated form
    key = [3, 1, 4, 1, 5, 9, 2, 6]
    a0, a1, z0, z1 = (rand_u64(rng, n) for _ in range(4))
    a0[:4], a1[:4] = np.array([5, 2**63, 0, 7], dtype=np.uint64), np.array([0, 0, 0, 2**64 - 7], dtype=np.uint64)
    ks = oracle.prg_fill(key, 99, 0, n)
    v = a0 + a1
    want = np.where(v.astype(np.int64) > 0, v, 0).astype(np.uint64) - ks
    assert np.array_equal(to_np(cgb.ideal_relu_reshare(key, 99, to_dev(a0), to_dev(a1))), want)
    gate = (z0 + z1).astype(np.int64) > 0
    want = np.where(gate, v, 0).astype(np.uint64) - ks
    assert np.array_equal(to_np(cgb.ideal_relu_reshare(key, 99, to_dev(a0), to_dev(a1), to_dev(z0), to_dev(z1))), want)


def test_prg_fill_multi_and_rowmul_sub_match_oracle(cgb, oracle):
    key = [7, 7, 7, 1, 2, 3, 4, 5]
    rng = np.random.default_rng(21)
    sizes = [1, 8, 9, 1433 * 16, 100_003, 7, 64, 5000, 12, 300, 4096, 33, 2, 77777, 640, 19, 250]  # 17 segments: two launches
    outs = [cgb.empty(n) for n in sizes]
    outs_b = [cgb.empty(n) if i % 3 == 0 else None for i, n in enumerate(sizes)]
    segs = []
    for i, n in enumerate(sizes):
        if i % 2 == 0:
            segs.append((outs[i], 1000 + i, 2000 + i, outs_b[i]))
        else:
            segs.append((outs[i], 1000 + i))
    before = cgb.launches
    cgb.prg_fill_multi(key, segs)
    assert cgb.launches - before == 2
    for i, n in enumerate(sizes):
        a = oracle.prg_fill(key, 1000 + i, 0, n)
        if i % 2 == 0:
            b = oracle.prg_fill(key, 2000 + i, 0, n)
            assert np.array_equal(to_np(outs[i]), a + b), i
            if outs_b[i] is not None:
                assert np.array_equal(to_np(outs_b[i]), b), i
        else:
            assert np.array_equal(to_np(outs[i]), a), i
    # an unaligned destination (a block inside a larger matrix starting at an odd word)
    big = cgb.empty(1001)
    cgb.prg_fill_multi(key, [(big[1:], 55)])
    assert np.array_equal(to_np(big[1:]), oracle.prg_fill(key, 55, 0, 1000))
    rows, D = 1354, 7
    a, c = rand_u64(rng, rows, D), rand_u64(rng, rows, D)
    b = rand_u64(rng, rows)
    assert np.array_equal(to_np(cgb.rowmul_sub(to_dev(a), to_dev(b), to_dev(c))), a * b[:, None] - c)


@pytest.mark.parametrize("n,C", [(1, 3), (255, 7), (256, 7), (1354, 7), (21168, 40)])
def test_prediction_metrics_match_host_loop(cgb, n, C):
    """gcn.h:603-632 as the engine's host loop computed it until round 2: zeros -> 0.001, first maximal class, -log p[label]."""
    import torch

    rng = np.random.default_rng(n + C)
    f = 16
    p = rng.random((n, C))
    p /= p.sum(1, keepdims=True)
    fixed = np.floor(p * (1 << f)).astype(np.int64)
    fixed[rng.integers(0, n, size=max(1, n // 10)), rng.integers(0, C, size=max(1, n // 10))] = 0  # exact zeros
    if n > 5:
        fixed[3] = fixed[3, 0]  # a tie over the whole row: the first class wins
    s1 = rand_u64(rng, n, C)
    s0 = fixed.view(np.uint64) - s1
    labels = rng.integers(0, C, size=n).astype(np.int32)
    train, val = int(n * 0.4), int(n * 0.2)
    pd = fixed.astype(np.float64) / (1 << f)
    pd[pd == 0] = 0.001
    arg = pd.argmax(1)  # first maximum
    ok = arg == labels
    want = (float(-np.log(np.maximum(pd[np.arange(n), labels], 1e-30)).sum()), float(ok.sum()), float(ok[:train].sum()),
            float(ok[train + val:].sum()))
    got = cgb.prediction_metrics(to_dev(s0), to_dev(s1), torch.from_numpy(labels).cuda(), train, val, f)
    assert got[1:] == want[1:]
    assert abs(got[0] - want[0]) <= 1e-9 * max(1.0, abs(want[0]))
