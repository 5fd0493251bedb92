"""Property tests (hypothesis) of the C oracle against exact Python integers on small random shapes: ring matmul with transA /
accumulate, the Beaver recombination, the SecureML truncation pair, and the fused gather against a plain double loop."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import pyoracle as po

MASK = (1 << 64) - 1
u64s = st.integers(min_value=0, max_value=MASK)


def mat(draw, r, c):
    return np.array(draw(st.lists(u64s, min_size=r * c, max_size=r * c)), dtype=np.uint64).reshape(r, c)


@settings(max_examples=60, deadline=None)
@given(st.data())
def test_matmul_and_beaver_against_python_ints(data):
    po.build()
    M, K, N = (data.draw(st.integers(1, 5)) for _ in range(3))
    A, B, C0 = mat(data.draw, M, K), mat(data.draw, K, N), mat(data.draw, M, N)
    want = [[(int(C0[i, j]) + sum(int(A[i, k]) * int(B[k, j]) for k in range(K))) & MASK for j in range(N)] for i in range(M)]
    assert po.matmul(A, B, C_in=C0).tolist() == want
    assert po.matmul(np.ascontiguousarray(A.T), B, transA=True, C_in=C0).tolist() == want
    # Beaver: shares of X, W and of a triple (U, V, Z = U V) recombine to X W (before truncation: f < 0)
    X1, W1, U0, U1, V0, V1, Z0 = (mat(data.draw, *s) for s in ((M, K), (K, N), (M, K), (M, K), (K, N), (K, N), (M, N)))
    X0, W0 = A - X1, B - W1
    Z1 = po.matmul(U0 + U1, V0 + V1) - Z0
    E, F = (X0 - U0) + (X1 - U1), (W0 - V0) + (W1 - V1)
    c0 = po.beaver_matmul_finish(E, F, U0, V0, Z0, 0, -1)
    c1 = po.beaver_matmul_finish(E, F, U1, V1, Z1, 1, -1)
    assert (c0 + c1).tolist() == [[sum(int(A[i, k]) * int(B[k, j]) for k in range(K)) & MASK for j in range(N)] for i in range(M)]


@settings(max_examples=100, deadline=None)
@given(st.integers(-(1 << 40), 1 << 40), u64s, st.integers(1, 30))
def test_local_truncation_pair_is_within_one_ulp(x, r, f):
    """trunc_0(z0) + trunc_1(z1) = floor(x / 2^f) + {0, +-1} for |x| far below 2^63, whatever the mask r is -- unless the
    shares wrap, which needs r within |x| of 0 or 2^64 (probability |x| / 2^63, excluded here)."""
    po.build()
    z0, z1 = (x - r) & MASK, r
    if r < (1 << 41) or r > MASK - (1 << 41):
        return
    lib = po.lib()
    got = (lib.orc_trunc_share(z0, f, 0) + lib.orc_trunc_share(z1, f, 1)) & MASK
    got = got - (1 << 64) if got >> 63 else got
    assert abs(got - (x >> f)) <= 1


@settings(max_examples=40, deadline=None)
@given(st.data())
def test_gather_sum_against_double_loop(data):
    po.build()
    n_dst, n_src, D = data.draw(st.integers(1, 6)), data.draw(st.integers(1, 6)), data.draw(st.integers(1, 4))
    degs = data.draw(st.lists(st.integers(0, 4), min_size=n_dst, max_size=n_dst))
    rowptr = np.zeros(n_dst + 1, dtype=np.uint32)
    rowptr[1:] = np.cumsum(degs)
    col = np.array(data.draw(st.lists(st.integers(0, n_src - 1), min_size=int(rowptr[-1]), max_size=int(rowptr[-1]))), dtype=np.uint32)
    x, delta = mat(data.draw, n_src, D), mat(data.draw, n_dst, D)
    want = [[(int(delta[v, j]) + sum(int(x[col[e], j]) for e in range(rowptr[v], rowptr[v + 1]))) & MASK for j in range(D)]
            for v in range(n_dst)]
    assert po.gather_sum_csr(rowptr, col, x, delta).tolist() == want
