"""Pins the oracle's ring arithmetic against independent numpy / Python big-int restatements."""
import numpy as np

MASK = (1 << 64) - 1
rng = np.random.default_rng(1234)


def rand_u64(*shape):
    return rng.integers(0, 1 << 64, size=shape, dtype=np.uint64)


def to_int(a):
    return [[int(v) for v in row] for row in a]


def test_encode_decode_truncation_toward_zero(oracle):
    f = 16
    L = oracle.lib()
    assert L.orc_encode_fixed(1.0, f) == 1 << 16
    assert L.orc_encode_fixed(0.5, f) == 1 << 15
    # gcn.h:191 pattern static_cast<uint64_t>(x * (1<<f)): fractions beyond f bits are dropped toward zero
    assert L.orc_encode_fixed(1.0 / 3.0, f) == int((1.0 / 3.0) * 65536)
    assert L.orc_encode_fixed(-1.0 / 3.0, f) == (-int((1.0 / 3.0) * 65536)) & MASK
    assert L.orc_decode_fixed((-(3 << 15)) & MASK, f) == -1.5
    x = rng.normal(size=1000) * 100
    assert np.allclose(oracle.decode(oracle.encode(x, f), f), x, atol=2.0 ** -f)


def test_share_split_reconstructs(oracle):
    f = 16
    x = rng.normal(size=(37, 5))
    key = [1, 2, 3, 4, 5, 6, 7, 8]
    s0, s1 = oracle.share_split(x, f, key, stream=9, word_offset=11)
    assert np.array_equal(s1.ravel(), oracle.prg_fill(key, 9, 11, x.size))
    assert np.array_equal(s0 + s1, oracle.encode(x, f))
    assert np.allclose(oracle.open_decode(s0, s1, f), x, atol=2.0 ** -f)


def test_local_truncation_error_at_most_one_ulp(oracle):
    f = 16
    vals = rng.integers(-(1 << 40), 1 << 40, size=5000, dtype=np.int64)
    z = vals.astype(np.uint64)
    r = rand_u64(z.size)
    z0, z1 = z - r, r
    t = oracle.trunc(z0, f, 0) + oracle.trunc(z1, f, 1)
    got = t.astype(np.int64)
    want = vals >> f  # floor
    assert np.all(np.abs(got - want) <= 1)
    # scalar definition
    L = oracle.lib()
    assert L.orc_trunc_share(0xFFFF_0000_0000_1234, 16, 0) == 0xFFFF_0000_0000_1234 >> 16
    assert L.orc_trunc_share(5, 16, 1) == (-(((-5) & MASK) >> 16)) & MASK


def test_matmul_against_bigints(oracle):
    A, B = rand_u64(7, 13), rand_u64(13, 5)
    want = [[sum(int(A[i, k]) * int(B[k, j]) for k in range(13)) & MASK for j in range(5)] for i in range(7)]
    assert to_int(oracle.matmul(A, B)) == want
    assert to_int(oracle.matmul(np.ascontiguousarray(A.T), B, transA=True)) == want
    C0 = rand_u64(7, 5)
    acc = oracle.matmul(A, B, C_in=C0)
    assert to_int(acc) == [[(want[i][j] + int(C0[i, j])) & MASK for j in range(5)] for i in range(7)]


def test_beaver_matmul_reconstructs_product(oracle):
    f = 16
    M, K, N = 9, 6, 4
    X = oracle.encode(rng.normal(size=(M, K)), f)
    W = oracle.encode(rng.normal(size=(K, N)), f)
    X1, W1 = rand_u64(M, K), rand_u64(K, N)
    X0, W0 = X - X1, W - W1
    U0, U1, V0, V1, Z0 = rand_u64(M, K), rand_u64(M, K), rand_u64(K, N), rand_u64(K, N), rand_u64(M, N)
    Z1 = oracle.matmul(U0 + U1, V0 + V1) - Z0
    E = (X0 - U0) + (X1 - U1)
    F = (W0 - V0) + (W1 - V1)
    # without truncation the shares add up to X*W exactly
    C0 = oracle.beaver_matmul_finish(E, F, U0, V0, Z0, 0, -1)
    C1 = oracle.beaver_matmul_finish(E, F, U1, V1, Z1, 1, -1)
    assert np.array_equal(C0 + C1, oracle.matmul(X, W))
    # with truncation: within one ulp of the floor of the exact product
    C0 = oracle.beaver_matmul_finish(E, F, U0, V0, Z0, 0, f)
    C1 = oracle.beaver_matmul_finish(E, F, U1, V1, Z1, 1, f)
    exact = oracle.matmul(X, W).astype(np.int64) >> f
    assert np.all(np.abs((C0 + C1).astype(np.int64) - exact) <= 1)
    want = oracle.decode(X, f) @ oracle.decode(W, f)
    assert np.allclose(oracle.decode(C0 + C1, f), want, atol=1e-3)


def test_rowmul_beaver_and_mux(oracle):
    f = 16
    rows, D = 11, 6
    x = oracle.encode(rng.normal(size=(rows, D)), f)
    s = oracle.encode(rng.uniform(0.1, 1.0, size=rows), f)
    x1, s1 = rand_u64(rows, D), rand_u64(rows)
    x0, s0 = x - x1, s - s1
    a0, a1, b0, b1, c0 = rand_u64(rows, D), rand_u64(rows, D), rand_u64(rows), rand_u64(rows), rand_u64(rows, D)
    c1 = (a0 + a1) * (b0 + b1)[:, None] - c0
    e = (x0 - a0) + (x1 - a1)
    fv = (s0 - b0) + (s1 - b1)
    y0 = oracle.rowmul_beaver_finish(e, fv, a0, b0, c0, 0, -1)
    y1 = oracle.rowmul_beaver_finish(e, fv, a1, b1, c1, 1, -1)
    assert np.array_equal(y0 + y1, x * s[:, None])
    y0 = oracle.rowmul_beaver_finish(e, fv, a0, b0, c0, 0, f)
    y1 = oracle.rowmul_beaver_finish(e, fv, a1, b1, c1, 1, f)
    assert np.allclose(oracle.decode(y0 + y1, f), oracle.decode(x, f) * oracle.decode(s, f)[:, None], atol=1e-3)


def test_public_scale_and_apply_gradient(oracle):
    f = 16
    W = oracle.encode(rng.normal(size=50), f)
    d = oracle.encode(rng.normal(size=50), f)
    r = rand_u64(50)
    lr = int(0.5 * (1 << f))
    new0 = oracle.apply_gradient(W - r, d - r, lr, f, 0)
    new1 = oracle.apply_gradient(r, r, lr, f, 1)
    assert np.allclose(oracle.decode(new0 + new1, f), oracle.decode(W, f) - 0.5 * oracle.decode(d, f), atol=1e-3)
    c = int((1.0 / 7) * (1 << f))
    s0, s1 = oracle.scale_public(W - r, c, f, 0), oracle.scale_public(r, c, f, 1)
    assert np.allclose(oracle.decode(s0 + s1, f), oracle.decode(W, f) * (c / 65536.0), atol=1e-3)
    assert np.array_equal(oracle.scale_public(W, c, f, 0), (W * np.uint64(c)) >> np.uint64(f))


def test_gather_expand_segsum(oracle):
    n_src, n_dst, D = 40, 23, 5
    deg = rng.integers(0, 9, size=n_dst)
    deg[3] = 0
    rowptr = np.zeros(n_dst + 1, dtype=np.uint32)
    rowptr[1:] = np.cumsum(deg)
    col = rng.integers(0, n_src, size=int(rowptr[-1])).astype(np.uint32)
    x, delta = rand_u64(n_src, D), rand_u64(n_dst, D)
    want = delta.copy()
    for v in range(n_dst):
        for e in range(rowptr[v], rowptr[v + 1]):
            want[v] += x[col[e]]
    assert np.array_equal(oracle.gather_sum_csr(rowptr, col, x, delta), want)
    assert np.array_equal(oracle.gather_sum_csr(rowptr, col, x), want - delta)
    # expand -> segsum(dup) -> first-of-group extract == fused gather (the reference's four-step dataflow)
    exp = oracle.expand_rows(col, x)
    assert np.array_equal(exp, x[col])
    dup = oracle.segsum(rowptr, exp, dup=True)
    compact = oracle.segsum(rowptr, exp, dup=False)
    assert np.array_equal(compact, want - delta)
    for v in range(n_dst):
        for e in range(rowptr[v], rowptr[v + 1]):
            assert np.array_equal(dup[e], compact[v])
    # allowMissing
    idx = np.array([0, oracle.NO_ROW, 5], dtype=np.uint32)
    d3 = rand_u64(3, D)
    got = oracle.expand_rows(idx, x, d3)
    assert np.array_equal(got[1], d3[1]) and np.array_equal(got[2], x[5] + d3[2])


def test_transpose_cond_add(oracle):
    x = rand_u64(6, 9)
    assert np.array_equal(oracle.transpose(x), x.T)
    v, u = rand_u64(6, 9), rand_u64(6, 9)
    cond = np.array([1, 0, 1, 1, 0, 0], dtype=np.uint8)
    assert np.array_equal(oracle.cond_add(v, u, cond), v + u * cond[:, None].astype(np.uint64))


def test_det_exp_c_equals_numpy_and_tracks_libm():
    """orc_det_exp (the exp of the 2PC-residual softmax stand-in) is plain IEEE arithmetic: the C build and the numpy
    restatement agree bit for bit, and both stay within 2 ulp of libm over the range a max-subtracted softmax visits."""
    import math

    from oracle import pyoracle as po

    po.build()
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.linspace(-60, 0, 6001), -rng.exponential(8, 5000), [-700.0, -700.5, -745.0, -1e4, 0.0, -1e-300]])
    c = np.array([po.det_exp(v) for v in xs])
    n = po.det_exp_numpy(xs)
    assert np.array_equal(c.view(np.uint64), n.view(np.uint64))
    ref = np.array([math.exp(v) if v > -700 else 0.0 for v in xs])
    ok = ref > 1e-290
    assert np.max(np.abs(c[ok] - ref[ok]) / ref[ok]) < 5e-16
    assert (c[xs < -700] == 0).all() and po.det_exp(0.0) == 1.0


def test_ideal_softmax_c_equals_numpy():
    from oracle import pyoracle as po

    rng = np.random.default_rng(1)
    for n, C, f in [(1, 3, 16), (400, 7, 16), (90, 40, 13)]:
        z = (rng.normal(0, 4, size=(n, C)) * (1 << f)).astype(np.int64).view(np.uint64)
        z1 = rng.integers(0, 1 << 64, size=(n, C), dtype=np.uint64)
        labels = rng.integers(0, C, size=n)
        a = po.ideal_softmax(z - z1, z1, labels, n // 2, f)
        b = po.ideal_softmax_numpy(z - z1, z1, labels, n // 2, f)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        p = a[0].astype(np.int64) / float(1 << f)
        zz = z.view(np.int64) / float(1 << f)
        e = np.exp(zz - zz.max(axis=1, keepdims=True))
        assert np.abs(p - e / e.sum(axis=1, keepdims=True)).max() < 2.0 / (1 << f)


def test_softmax_golden_vectors():
    """tests/golden/ideal_softmax.json (made by make_softmax_golden.py): the C oracle and its numpy restatement reproduce the
    committed bits of the restated exp and of the prediction-layer stand-in."""
    import json
    import os

    from oracle import pyoracle as po

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ideal_softmax.json")))
    for x, bits in g["det_exp"].items():
        assert int(np.float64(po.det_exp(float(x))).view(np.uint64)) == bits, x
        assert int(po.det_exp_numpy(np.array([float(x)]))[0].view(np.uint64)) == bits, x
    shape = tuple(g["shape"])
    z0 = np.array(g["z0"], dtype=np.uint64).reshape(shape)
    z1 = np.array(g["z1"], dtype=np.uint64).reshape(shape)
    for fn in (po.ideal_softmax, po.ideal_softmax_numpy):
        P, pmy = fn(z0, z1, np.array(g["labels"]), g["train_rows"], g["f"])
        assert P.ravel().tolist() == g["P"] and pmy.ravel().tolist() == g["pmy"], fn.__name__
