"""ctypes/numpy front end of the CPU oracle (oracle/libcgb_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Parity unpinned by the reference (see cgb_oracle.h); pinned by tests/test_oracle_*.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcgb_oracle.so")
NO_ROW = 0xFFFFFFFF
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("cgb_oracle.c", "cgb_oracle.h", "Makefile")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "libcgb_oracle.so"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_encode_fixed.restype = C.c_uint64
        _lib.orc_encode_fixed.argtypes = [C.c_double, C.c_int]
        _lib.orc_decode_fixed.restype = C.c_double
        _lib.orc_decode_fixed.argtypes = [C.c_uint64, C.c_int]
        _lib.orc_trunc_share.restype = C.c_uint64
        _lib.orc_trunc_share.argtypes = [C.c_uint64, C.c_int, C.c_int]
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_det_exp.restype = C.c_double
        _lib.orc_det_exp.argtypes = [C.c_double]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u64(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint64, a.dtype
    return a


def _u32(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint32, a.dtype
    return a


def _key(key):
    return (C.c_uint32 * 8)(*[int(k) & 0xFFFFFFFF for k in key])


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(n))


def chacha20_block(key, counter, nonce):
    out = np.zeros(16, dtype=np.uint32)
    lib().orc_chacha20_block(_key(key), C.c_uint32(counter), (C.c_uint32 * 3)(*nonce), _p(out))
    return out


def prg_fill(key, stream, word_offset, n_words):
    out = np.zeros(n_words, dtype=np.uint64)
    lib().orc_prg_fill(_key(key), C.c_uint64(stream), C.c_uint64(word_offset), _p(out), C.c_size_t(n_words))
    return out


def encode(x, f):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(x.shape, dtype=np.uint64)
    lib().orc_encode(_p(x), _p(out), C.c_size_t(x.size), C.c_int(f))
    return out


def decode(v, f):
    v = _u64(v)
    out = np.zeros(v.shape, dtype=np.float64)
    lib().orc_decode(_p(v), _p(out), C.c_size_t(v.size), C.c_int(f))
    return out


def share_split(x, f, key, stream, word_offset=0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s0 = np.zeros(x.shape, dtype=np.uint64)
    s1 = np.zeros(x.shape, dtype=np.uint64)
    lib().orc_share_split(_p(x), C.c_size_t(x.size), C.c_int(f), _key(key), C.c_uint64(stream),
                          C.c_uint64(word_offset), _p(s0), _p(s1))
    return s0, s1


def open_decode(s0, s1, f):
    s0, s1 = _u64(s0), _u64(s1)
    out = np.zeros(s0.shape, dtype=np.float64)
    lib().orc_open_decode(_p(s0), _p(s1), _p(out), C.c_size_t(s0.size), C.c_int(f))
    return out


def add(a, b):
    a, b = _u64(a), _u64(b)
    out = np.zeros(a.shape, dtype=np.uint64)
    lib().orc_add(_p(a), _p(b), _p(out), C.c_size_t(a.size))
    return out


def sub(a, b):
    a, b = _u64(a), _u64(b)
    out = np.zeros(a.shape, dtype=np.uint64)
    lib().orc_sub(_p(a), _p(b), _p(out), C.c_size_t(a.size))
    return out


def trunc(x, f, share):
    x = _u64(x)
    out = np.zeros(x.shape, dtype=np.uint64)
    lib().orc_trunc(_p(x), _p(out), C.c_size_t(x.size), C.c_int(f), C.c_int(share))
    return out


def scale_public(x, c, f, share):
    x = _u64(x)
    out = np.zeros(x.shape, dtype=np.uint64)
    lib().orc_scale_public(_p(x), C.c_uint64(c & (2**64 - 1)), _p(out), C.c_size_t(x.size), C.c_int(f), C.c_int(share))
    return out


def apply_gradient(W, d, lr, f, share):
    W, d = _u64(W), _u64(d)
    out = np.zeros(W.shape, dtype=np.uint64)
    lib().orc_apply_gradient(_p(W), _p(d), C.c_uint64(lr & (2**64 - 1)), _p(out), C.c_size_t(W.size), C.c_int(f),
                             C.c_int(share))
    return out


def rowmul_beaver_finish(e, fv, a, b, c, share, f):
    e, fv, a, b, c = map(_u64, (e, fv, a, b, c))
    rows, D = e.shape
    out = np.zeros(e.shape, dtype=np.uint64)
    lib().orc_rowmul_beaver_finish(_p(e), _p(fv), _p(a), _p(b), _p(c), _p(out), C.c_size_t(rows), C.c_size_t(D),
                                   C.c_int(share), C.c_int(f))
    return out


def cond_add(v, u, cond):
    v, u = _u64(v), _u64(u)
    cond = np.ascontiguousarray(cond, dtype=np.uint8)
    rows, D = v.shape
    out = np.zeros(v.shape, dtype=np.uint64)
    lib().orc_cond_add(_p(v), _p(u), _p(cond), _p(out), C.c_size_t(rows), C.c_size_t(D))
    return out


def det_exp(x):
    return float(lib().orc_det_exp(float(x)))


_EXP_C = [float.fromhex(h) for h in (
    "0x1.0000000000000p+0", "0x1.0000000000000p+0", "0x1.0000000000000p-1", "0x1.5555555555555p-3", "0x1.5555555555555p-5",
    "0x1.1111111111111p-7", "0x1.6c16c16c16c17p-10", "0x1.a01a01a01a01ap-13", "0x1.a01a01a01a01ap-16", "0x1.71de3a556c734p-19",
    "0x1.27e4fb7789f5cp-22", "0x1.ae64567f544e4p-26", "0x1.1eed8eff8d898p-29", "0x1.6124613a86d09p-33")]


def det_exp_numpy(x):
    """numpy restatement of orc_det_exp (every ufunc rounds once, so no fused multiply-add can sneak in)."""
    x = np.asarray(x, dtype=np.float64)
    zero = x < -700.0
    xc = np.minimum(np.where(zero, 0.0, x), 700.0)
    k = np.floor(xc * float.fromhex("0x1.71547652b82fep+0") + 0.5)
    r = (xc - k * float.fromhex("0x1.62e42fee00000p-1")) - k * float.fromhex("0x1.a39ef35793c76p-33")
    p = np.full_like(xc, _EXP_C[13])
    for i in range(12, -1, -1):
        p = p * r + _EXP_C[i]
    two_k = ((k.astype(np.int64) + 1023).astype(np.uint64) << np.uint64(52)).view(np.float64)
    return np.where(zero, 0.0, p * two_k)


def ideal_softmax(z0, z1, labels, train_rows, f):
    """2PC-RESIDUAL stand-in for the prediction layer (gcn.h:578,591) on reconstructed logits: returns (P, P - onehot)."""
    z0, z1 = _u64(z0), _u64(z1)
    n, Cc = z0.shape
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    P = np.zeros((n, Cc), dtype=np.uint64)
    pmy = np.zeros((n, Cc), dtype=np.uint64)
    lib().orc_ideal_softmax(_p(z0), _p(z1), _p(lab), C.c_size_t(n), C.c_size_t(Cc), C.c_size_t(train_rows), f, _p(P), _p(pmy))
    return P, pmy


def ideal_softmax_numpy(z0, z1, labels, train_rows, f):
    """numpy restatement of orc_ideal_softmax (left-to-right row sums)."""
    z = (_u64(z0) + _u64(z1)).view(np.int64).astype(np.float64) / float(1 << f)
    n, Cc = z.shape
    e = det_exp_numpy(z - z.max(axis=1, keepdims=True))
    tot = np.zeros(n)
    for j in range(Cc):
        tot = tot + e[:, j]
    P = ((e / tot[:, None]) * float(1 << f)).astype(np.int64).view(np.uint64)
    onehot = np.zeros((n, Cc), dtype=np.uint64)
    onehot[np.arange(n), np.asarray(labels)] = np.uint64(1 << f)
    pmy = P - onehot
    pmy[train_rows:] = 0
    return P, pmy


def transpose(x):
    x = _u64(x)
    rows, cols = x.shape
    out = np.zeros((cols, rows), dtype=np.uint64)
    lib().orc_transpose(_p(x), _p(out), C.c_size_t(rows), C.c_size_t(cols))
    return out


def gather_sum_csr(rowptr, col, x, delta=None, out=None):
    """y = delta + A x.  `out` (n_rows x D uint64, C-contiguous) avoids the allocation; with delta is out the sum is
    accumulated in place."""
    rowptr, col, x = _u32(rowptr), _u32(col), _u64(x)
    n_rows = rowptr.size - 1
    D = x.shape[1]
    if delta is not None:
        delta = _u64(delta)
    if out is None:
        y = np.zeros((n_rows, D), dtype=np.uint64)
    else:
        y = out
        assert y.dtype == np.uint64 and y.flags.c_contiguous and y.shape == (n_rows, D)
    lib().orc_gather_sum_csr(_p(rowptr), _p(col), _p(x), _p(delta), _p(y), C.c_size_t(n_rows), C.c_size_t(D))
    return y


def expand_rows(idx, x, delta=None):
    idx, x = _u32(idx), _u64(x)
    D = x.shape[1]
    if delta is not None:
        delta = _u64(delta)
    y = np.zeros((idx.size, D), dtype=np.uint64)
    lib().orc_expand_rows(_p(idx), C.c_size_t(idx.size), _p(x), _p(delta), _p(y), C.c_size_t(D))
    return y


def segsum(segptr, inp, dup):
    segptr, inp = _u32(segptr), _u64(inp)
    n_seg = segptr.size - 1
    D = inp.shape[1]
    out = np.zeros((inp.shape[0] if dup else n_seg, D), dtype=np.uint64)
    lib().orc_segsum(_p(segptr), C.c_size_t(n_seg), _p(inp), _p(out), C.c_size_t(D), C.c_int(1 if dup else 0))
    return out


def matmul(A, B, transA=False, C_in=None):
    A, B = _u64(A), _u64(B)
    if transA:
        K, M = A.shape
    else:
        M, K = A.shape
    N = B.shape[1]
    assert B.shape[0] == K
    out = np.zeros((M, N), dtype=np.uint64) if C_in is None else _u64(C_in).copy()
    lib().orc_matmul(_p(A), _p(B), _p(out), C.c_size_t(M), C.c_size_t(K), C.c_size_t(N), C.c_int(int(transA)),
                     C.c_int(0 if C_in is None else 1))
    return out


def beaver_matmul_finish(E, F, U, V, Z, share, f):
    E, F, U, V, Z = map(_u64, (E, F, U, V, Z))
    M, K = E.shape
    N = F.shape[1]
    out = np.zeros((M, N), dtype=np.uint64)
    lib().orc_beaver_matmul_finish(_p(E), _p(F), _p(U), _p(V), _p(Z), _p(out), C.c_size_t(M), C.c_size_t(K),
                                   C.c_size_t(N), C.c_int(share), C.c_int(f))
    return out
