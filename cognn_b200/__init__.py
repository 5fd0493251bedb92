"""cognn_b200 -- B200-native share-local engine for CoGNN's secret-shared GCN path.

The product is libcognn_b200.so (hand-written sm_100a CUDA behind the C ABI of include/cognn_b200.h) plus the C++
host engine in cognn_b200/host.  This Python package is plumbing for tests and bench.py: it binds the C ABI with
ctypes and lets torch own device buffers / streams / process groups.  u64 shares are stored in torch.int64 tensors
(same bits).
"""
import ctypes as C

from . import _lib
from ._lib import CgbError, LIB_PATH, SIGNATURES, load  # noqa: F401

SCALER_BITS = 16
NO_ROW = 0xFFFFFFFF


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Csr:
    def __init__(self, ctx, handle, n_rows, n_edges, n_src_rows):
        self.ctx, self.handle = ctx, handle
        self.n_rows, self.n_edges, self.n_src_rows = n_rows, n_edges, n_src_rows

    @property
    def n_nonempty(self):
        return int(self.ctx.lib.cgb_csr_num_nonempty_rows(self.handle))

    def nonempty_rows(self):
        """Device int32 tensor (a copy) of the destination rows that have at least one edge, ascending."""
        t = self.ctx.torch
        n = self.n_nonempty
        out = t.empty(n, dtype=t.int32, device=self.ctx._dev())
        if n:
            self.ctx.check(self.ctx.lib.cgb_d2d(self.ctx.handle, C.c_void_p(out.data_ptr()),
                                                C.c_void_p(self.ctx.lib.cgb_csr_nonempty_rows(self.handle)), 4 * n))
        return out

    def destroy(self):
        if self.handle is not None:
            self.ctx.lib.cgb_csr_destroy(self.ctx.handle, self.handle)
            self.handle = None


class Context:
    """One cgb_ctx bound to a CUDA device and (by default) torch's current stream."""

    def __init__(self, device=0, own_stream=False):
        import torch

        self.torch = torch
        self.lib = load()
        self.device = int(device)
        h = C.c_void_p()
        if own_stream:
            rc = self.lib.cgb_ctx_create(self.device, C.byref(h))
        else:
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream().cuda_stream
            rc = self.lib.cgb_ctx_create_on_stream(self.device, C.c_void_p(stream), C.byref(h))
        if rc != 0:
            raise CgbError(f"cgb_ctx_create failed ({rc}): {self.lib.cgb_last_error(None).decode()}")
        self.handle = h

    # -- helpers ---------------------------------------------------------------------------------------------
    def check(self, rc):
        if rc != 0:
            raise CgbError(f"cgb error {rc}: {self.lib.cgb_last_error(self.handle).decode()}")

    def close(self):
        if self.handle is not None:
            self.lib.cgb_ctx_destroy(self.handle)
            self.handle = None

    def sync(self):
        self.check(self.lib.cgb_ctx_sync(self.handle))

    @property
    def launches(self):
        return int(self.lib.cgb_ctx_launch_count(self.handle))

    def _dev(self):
        return self.torch.device("cuda", self.device)

    def empty(self, *shape):
        return self.torch.empty(*shape, dtype=self.torch.int64, device=self._dev())

    def _u64(self, t):
        assert t.dtype == self.torch.int64 and t.is_cuda and t.is_contiguous(), "expect contiguous cuda int64"
        return t

    # -- (1) gather --------------------------------------------------------------------------------------------
    def csr_create(self, rowptr, col, n_src_rows):
        """rowptr/col: int32 torch tensors (cuda -> device path, cpu -> host path)."""
        t = self.torch
        n_rows = rowptr.numel() - 1
        n_edges = col.numel()
        h = C.c_void_p()
        assert rowptr.dtype == t.int32 and col.dtype == t.int32
        rowptr, col = rowptr.contiguous(), col.contiguous()
        fn = self.lib.cgb_csr_create_device if rowptr.is_cuda else self.lib.cgb_csr_create
        self.check(fn(self.handle, _ptr(rowptr), _ptr(col), n_rows, n_edges, n_src_rows, C.byref(h)))
        return Csr(self, h, n_rows, n_edges, n_src_rows)

    def gather_sum(self, csr, x, delta=None, out=None):
        D = x.shape[1]
        assert x.shape[0] == csr.n_src_rows
        if out is None:
            out = self.empty(csr.n_rows, D)
        self.check(self.lib.cgb_gather_sum(self.handle, csr.handle, _ptr(self._u64(x)),
                                           _ptr(delta), _ptr(self._u64(out)), D))
        return out

    def gather_sum_compact(self, csr, x, delta=None, out=None, out_ptr=None):
        """Compact mirror-update block: row k = k-th non-empty destination row (csr.nonempty_rows()).  `out_ptr`: raw device
        address (e.g. an IPC-exported staging buffer) instead of a tensor."""
        D = x.shape[1]
        assert x.shape[0] == csr.n_src_rows
        if out_ptr is None:
            if out is None:
                out = self.empty(csr.n_nonempty, D)
            out_ptr = self._u64(out).data_ptr()
        self.check(self.lib.cgb_gather_sum_compact(self.handle, csr.handle, _ptr(self._u64(x)), _ptr(delta),
                                                   C.c_void_p(int(out_ptr)), D))
        return out

    def gather_sum_signal(self, csr, x, offsets, bases, compact, flags, value):
        """One launch; block b = rows [offsets[b], offsets[b+1]) -> bases[b] (raw address; dense or compact[b]), flags[b] (raw
        address or 0) raised with `value` when the block is complete."""
        n = len(bases)
        D = x.shape[1]
        offs = (C.c_uint32 * (n + 1))(*[int(o) for o in offsets])
        bs = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in bases])
        cm = (C.c_uint8 * n)(*[1 if c else 0 for c in compact])
        fl = (C.c_void_p * n)(*[C.c_void_p(int(f)) if f else None for f in flags])
        self.check(self.lib.cgb_gather_sum_signal(self.handle, csr.handle, _ptr(self._u64(x)), D, n, offs, bs, cm, fl,
                                                  int(value) & 0xFFFFFFFF))

    def scatter_add_rows(self, idx, src, v, assign=False, n=None, D=None, n_ctas=0):
        """v[idx[k], :] (+)= src[k, :].  `src` may be a raw device address (int; e.g. a peer's staging buffer) with n given."""
        D = v.shape[1] if D is None else D
        if isinstance(src, int):
            src_ptr = src
        else:
            src_ptr, n = self._u64(src).data_ptr(), src.shape[0]
        assert idx.dtype == self.torch.int32 and idx.is_cuda and idx.numel() >= n
        self.check(self.lib.cgb_scatter_add_rows(self.handle, _ptr(idx), n, C.c_void_p(int(src_ptr)), _ptr(self._u64(v)), D,
                                                 int(assign), n_ctas))
        return v

    def flag_signal(self, flag_ptr, value):
        self.check(self.lib.cgb_flag_signal(self.handle, C.c_void_p(int(flag_ptr)), int(value) & 0xFFFFFFFF))

    def flag_wait(self, flag_ptr, value, mode=0, err_ptr=None):
        self.check(self.lib.cgb_flag_wait(self.handle, C.c_void_p(int(flag_ptr)), int(value) & 0xFFFFFFFF, int(mode),
                                          C.c_void_p(int(err_ptr)) if err_ptr else None))

    def peer_round(self, segs, links, slot_words, ctas_per_seg=0, err_ptr=None):
        """segs: [(src_ptr, dst_ptr, n_words, link)] (push segments first), links: [(seq, done, wait_flag, signal_flag, recv)]
        (raw device addresses)."""
        sa = (_lib.XSeg * max(len(segs), 1))(*[_lib.XSeg(int(a), int(b), int(n), int(l)) for a, b, n, l in segs])
        la = (_lib.XLink * max(len(links), 1))(*[_lib.XLink(int(a), int(b), int(c), int(d), int(r)) for a, b, c, d, r in links])
        self.check(self.lib.cgb_peer_round(self.handle, sa, len(segs), la, len(links), int(slot_words), int(ctas_per_seg),
                                           C.c_void_p(int(err_ptr)) if err_ptr else None))

    def set_matmul_impl(self, impl):
        """'auto' | 'imad' | 'tc' | None (environment)."""
        self.check(self.lib.cgb_ctx_set_matmul_impl(self.handle, {None: -1, "auto": 0, "imad": 1, "tc": 2}[impl]))

    def probe_imad_peak(self):
        v = C.c_double()
        self.check(self.lib.cgb_probe_imad_peak(self.handle, C.byref(v)))
        return v.value

    def probe_tensor_i8_peak(self):
        v = C.c_double()
        self.check(self.lib.cgb_probe_tensor_i8_peak(self.handle, C.byref(v)))
        return v.value

    @property
    def last_kernel(self):
        return (self.lib.cgb_ctx_last_kernel(self.handle) or b"").decode()

    def gather_sum_blocks(self, csr, x, block_ptrs, block_offsets, delta=None):
        """block_ptrs: list of raw device addresses (ints; may be peer memory), block_offsets: row offsets (len + 1)."""
        D = x.shape[1]
        n = len(block_ptrs)
        bases = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in block_ptrs])
        offs = (C.c_uint32 * (n + 1))(*[int(o) for o in block_offsets])
        self.check(self.lib.cgb_gather_sum_blocks(self.handle, csr.handle, _ptr(self._u64(x)), _ptr(delta), D, n, bases, offs))

    def malloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.lib.cgb_malloc(self.handle, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr):
        self.check(self.lib.cgb_free(self.handle, C.c_void_p(ptr)))

    def ipc_export(self, ptr):
        h = (C.c_char * 64)()
        self.check(self.lib.cgb_ipc_export(self.handle, C.c_void_p(ptr), h))
        return bytes(h)

    def ipc_open(self, handle):
        p = C.c_void_p()
        buf = (C.c_char * 64).from_buffer_copy(handle)
        self.check(self.lib.cgb_ipc_open(self.handle, buf, C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        self.check(self.lib.cgb_ipc_close(self.handle, C.c_void_p(ptr)))

    def party_graph_build(self, edges, tid, T, me):
        """Device ingest (include/cognn_b200.h, cgb_party_graph_build).  edges: (E, 2) int64, tid: (V,) int64; cuda tensors take
        the device entry point, cpu tensors / numpy arrays the host one.  Returns a dict of torch tensors (device) plus the
        handle-owning PartyGraph object under "handle"."""
        t = self.torch
        edges = t.as_tensor(edges, dtype=t.int64).reshape(-1, 2).contiguous()
        tid = t.as_tensor(tid, dtype=t.int64).contiguous()
        h = C.c_void_p()
        fn = self.lib.cgb_party_graph_build if edges.is_cuda else self.lib.cgb_party_graph_build_host
        assert edges.is_cuda == tid.is_cuda
        self.check(fn(self.handle, _ptr(edges) if edges.numel() else None, edges.shape[0], _ptr(tid), tid.numel(), T, me,
                      C.byref(h)))
        lib = self.lib
        n_local, n_rows = lib.cgb_party_graph_num_local(h), lib.cgb_party_graph_num_rows(h)
        m = lib.cgb_party_graph_num_out_edges(h)
        offsets = list((C.c_uint32 * (T + 1)).from_address(lib.cgb_party_graph_offsets(h)))
        dev = self._dev()

        def grab(ptr, n, dtype):
            if n == 0:
                return t.empty(0, dtype=dtype, device=dev)
            typestr = {t.int64: "<i8", t.int32: "<i4", t.uint8: "|u1"}[dtype]

            class _V:
                __cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}

            return t.as_tensor(_V(), device=dev).clone()

        out = {"n_local": n_local, "n_rows": n_rows, "n_out_edges": m, "offsets": offsets,
               "vids": grab(lib.cgb_party_graph_vids(h), n_local, t.int64),
               "in_deg_raw": grab(lib.cgb_party_graph_in_deg_raw(h), n_local, t.int64),
               "in_deg": grab(lib.cgb_party_graph_in_deg(h), n_local, t.int64),
               "is_border": grab(lib.cgb_party_graph_is_border(h), n_local, t.uint8),
               "rowptr": grab(lib.cgb_party_graph_rowptr(h), n_rows + 1, t.int32),
               "col": grab(lib.cgb_party_graph_col(h), m, t.int32)}
        ch = C.c_void_p()
        self.check(lib.cgb_party_graph_csr(self.handle, h, C.byref(ch)))
        out["csr"] = Csr(self, ch, n_rows, m, n_local)
        self.check(lib.cgb_party_graph_destroy(self.handle, h))
        return out

    def peer_copy(self, dst_ptr, src_ptr, nbytes, n_ctas=0):
        """SM-driven copy between raw device addresses (ints); either side may be peer memory."""
        self.check(self.lib.cgb_peer_copy(self.handle, C.c_void_p(int(dst_ptr)), C.c_void_p(int(src_ptr)), nbytes, n_ctas))

    def expand_rows(self, idx, x, delta=None, out=None):
        D = x.shape[1]
        n_out = idx.numel()
        assert idx.dtype == self.torch.int32 and idx.is_cuda
        if out is None:
            out = self.empty(n_out, D)
        self.check(self.lib.cgb_expand_rows(self.handle, _ptr(idx), n_out, _ptr(self._u64(x)), _ptr(delta),
                                            _ptr(out), D))
        return out

    def segsum(self, segptr, inp, dup):
        D = inp.shape[1]
        n_seg = segptr.numel() - 1
        assert segptr.dtype == self.torch.int32 and segptr.is_cuda
        out = self.empty(inp.shape[0] if dup else n_seg, D)
        self.check(self.lib.cgb_segsum(self.handle, _ptr(segptr), n_seg, inp.shape[0], _ptr(self._u64(inp)),
                                       _ptr(out), D, 1 if dup else 0))
        return out

    # -- (2) matmul --------------------------------------------------------------------------------------------
    def matmul(self, A, B, transA=False, out=None, accumulate=False):
        if transA:
            K, M = A.shape
        else:
            M, K = A.shape
        assert B.shape[0] == K
        N = B.shape[1]
        if out is None:
            assert not accumulate
            out = self.empty(M, N)
        self.check(self.lib.cgb_matmul(self.handle, _ptr(self._u64(A)), _ptr(self._u64(B)), _ptr(self._u64(out)),
                                       M, K, N, int(transA), int(accumulate)))
        return out

    def beaver_matmul_finish(self, E, F, U, V, Z, share, f=SCALER_BITS):
        M, K = E.shape
        N = F.shape[1]
        out = self.empty(M, N)
        self.check(self.lib.cgb_beaver_matmul_finish(self.handle, _ptr(self._u64(E)), _ptr(self._u64(F)),
                                                     _ptr(self._u64(U)), _ptr(self._u64(V)), _ptr(self._u64(Z)),
                                                     _ptr(out), M, K, N, share, f))
        return out

    def beaver_matmul_finish_open(self, mine, peer, U, V, Z, share, f=SCALER_BITS):
        """`mine` ([E_i | F_i] flat) is opened in place."""
        M, K = U.shape
        N = V.shape[1]
        out = self.empty(M, N)
        self.check(self.lib.cgb_beaver_matmul_finish_open(self.handle, _ptr(self._u64(mine)), _ptr(self._u64(peer)),
                                                          _ptr(self._u64(U)), _ptr(self._u64(V)), _ptr(self._u64(Z)),
                                                          _ptr(out), M, K, N, share, f))
        return out

    # -- (3) elementwise ---------------------------------------------------------------------------------------
    def add(self, a, b, out=None):
        out = self.torch.empty_like(a) if out is None else out
        self.check(self.lib.cgb_add(self.handle, _ptr(self._u64(a)), _ptr(self._u64(b)), _ptr(out), a.numel()))
        return out

    def sub(self, a, b, out=None):
        out = self.torch.empty_like(a) if out is None else out
        self.check(self.lib.cgb_sub(self.handle, _ptr(self._u64(a)), _ptr(self._u64(b)), _ptr(out), a.numel()))
        return out

    def sum_n(self, tensors, out=None):
        n = len(tensors)
        out = self.torch.empty_like(tensors[0]) if out is None else out
        ptrs = (C.c_void_p * n)(*[C.c_void_p(self._u64(t).data_ptr()) for t in tensors])
        self.check(self.lib.cgb_sum_n(self.handle, ptrs, n, _ptr(out), tensors[0].numel()))
        return out

    def trunc(self, x, share, f=SCALER_BITS, out=None):
        out = self.torch.empty_like(x) if out is None else out
        self.check(self.lib.cgb_trunc(self.handle, _ptr(self._u64(x)), _ptr(out), x.numel(), f, share))
        return out

    def scale_public(self, x, c, share, f=SCALER_BITS, out=None):
        out = self.torch.empty_like(x) if out is None else out
        self.check(self.lib.cgb_scale_public(self.handle, _ptr(self._u64(x)), C.c_uint64(c & (2**64 - 1)), _ptr(out),
                                             x.numel(), f, share))
        return out

    def apply_gradient(self, W, d, lr, share, f=SCALER_BITS, out=None):
        out = self.torch.empty_like(W) if out is None else out
        self.check(self.lib.cgb_apply_gradient(self.handle, _ptr(self._u64(W)), _ptr(self._u64(d)),
                                               C.c_uint64(lr & (2**64 - 1)), _ptr(out), W.numel(), f, share))
        return out

    def rowmul_beaver_finish(self, e, fv, a, b, c, share, f=SCALER_BITS, out=None):
        rows, D = e.shape
        out = self.torch.empty_like(e) if out is None else out
        self.check(self.lib.cgb_rowmul_beaver_finish(self.handle, _ptr(self._u64(e)), _ptr(self._u64(fv)),
                                                     _ptr(self._u64(a)), _ptr(self._u64(b)), _ptr(self._u64(c)),
                                                     _ptr(out), rows, D, share, f))
        return out

    def rowmul_beaver_finish_open(self, mine, peer, a, b, c, share, f=SCALER_BITS, out=None):
        """`mine`, `peer`: the two halves of the opening, flat [rows * D | rows] words each."""
        rows, D = a.shape
        out = self.torch.empty_like(a) if out is None else out
        self.check(self.lib.cgb_rowmul_beaver_finish_open(self.handle, _ptr(self._u64(mine)), _ptr(self._u64(peer)),
                                                          _ptr(self._u64(a)), _ptr(self._u64(b)), _ptr(self._u64(c)),
                                                          _ptr(out), rows, D, share, f))
        return out

    def sub_pair(self, a0, b0, a1, b1, out=None):
        """out = [a0 - b0 | a1 - b1] flat; a1 None means 0 - b1."""
        n0, n1 = a0.numel(), b1.numel()
        out = self.empty(n0 + n1) if out is None else out
        self.check(self.lib.cgb_sub_pair(self.handle, _ptr(self._u64(a0)), _ptr(self._u64(b0)), n0,
                                         _ptr(self._u64(a1)) if a1 is not None else None, _ptr(self._u64(b1)), n1, _ptr(out)))
        return out

    def scale_apply_gradient(self, W, d, gs, lr, share, f=SCALER_BITS):
        """In place: d <- trunc(d * gs), W <- W - trunc(d * lr)."""
        m = 2**64 - 1
        self.check(self.lib.cgb_scale_apply_gradient(self.handle, _ptr(self._u64(W)), _ptr(self._u64(d)), C.c_uint64(gs & m),
                                                     C.c_uint64(lr & m), _ptr(d), _ptr(W), W.numel(), f, share))
        return W

    def avg_public(self, tensors, c, outs, share, f=SCALER_BITS):
        ni, no = len(tensors), len(outs)
        ip = (C.c_void_p * ni)(*[C.c_void_p(self._u64(t).data_ptr()) for t in tensors])
        op = (C.c_void_p * no)(*[C.c_void_p(self._u64(t).data_ptr()) for t in outs])
        self.check(self.lib.cgb_avg_public(self.handle, ip, ni, C.c_uint64(c & (2**64 - 1)), op, no, tensors[0].numel(), f, share))

    def ideal_relu_reshare(self, key, stream, a0, a1, z0=None, z1=None, out=None):
        out = self.torch.empty_like(a0) if out is None else out
        self.check(self.lib.cgb_ideal_relu_reshare(self.handle, _lib.key_array(key), stream, _ptr(self._u64(a0)),
                                                   _ptr(self._u64(a1)), _ptr(z0), _ptr(z1), _ptr(out), a0.numel()))
        return out

    def copy_segments(self, dsts, srcs):
        """dsts[j][:] = srcs[j][:] for every j, one launch per 16 segments."""
        n = len(dsts)
        dp = (C.c_void_p * max(n, 1))(*[C.c_void_p(self._u64(t).data_ptr()) for t in dsts])
        sp = (C.c_void_p * max(n, 1))(*[C.c_void_p(self._u64(t).data_ptr()) for t in srcs])
        nw = (C.c_uint64 * max(n, 1))(*[t.numel() for t in srcs])
        self.check(self.lib.cgb_copy_segments(self.handle, dp, sp, nw, n))

    def cond_add(self, v, u, cond, out=None):
        rows, D = v.shape
        assert cond.dtype == self.torch.uint8 and cond.is_cuda
        out = self.torch.empty_like(v) if out is None else out
        self.check(self.lib.cgb_cond_add(self.handle, _ptr(self._u64(v)), _ptr(self._u64(u)), _ptr(cond), _ptr(out),
                                         rows, D))
        return out

    def transpose(self, x):
        rows, cols = x.shape
        out = self.empty(cols, rows)
        self.check(self.lib.cgb_transpose(self.handle, _ptr(self._u64(x)), _ptr(out), rows, cols))
        return out

    def encode(self, x, f=SCALER_BITS):
        assert x.dtype == self.torch.float64 and x.is_cuda and x.is_contiguous()
        out = self.torch.empty(x.shape, dtype=self.torch.int64, device=x.device)
        self.check(self.lib.cgb_encode(self.handle, _ptr(x), _ptr(out), x.numel(), f))
        return out

    def decode(self, v, f=SCALER_BITS):
        out = self.torch.empty(v.shape, dtype=self.torch.float64, device=v.device)
        self.check(self.lib.cgb_decode(self.handle, _ptr(self._u64(v)), _ptr(out), v.numel(), f))
        return out

    def share_split(self, x, key, stream, word_offset=0, f=SCALER_BITS):
        assert x.dtype == self.torch.float64 and x.is_cuda and x.is_contiguous()
        s0 = self.torch.empty(x.shape, dtype=self.torch.int64, device=x.device)
        s1 = self.torch.empty_like(s0)
        self.check(self.lib.cgb_share_split(self.handle, _ptr(x), x.numel(), f, _lib.key_array(key), stream,
                                            word_offset, _ptr(s0), _ptr(s1)))
        return s0, s1

    def open_decode(self, s0, s1, f=SCALER_BITS):
        out = self.torch.empty(s0.shape, dtype=self.torch.float64, device=s0.device)
        self.check(self.lib.cgb_open_decode(self.handle, _ptr(self._u64(s0)), _ptr(self._u64(s1)), _ptr(out),
                                            s0.numel(), f))
        return out

    # -- 2PC-residual stand-ins (ideal functionality, not secure) ------------------------------------------------
    def ideal_relu(self, a0, a1):
        out = self.torch.empty_like(a0)
        self.check(self.lib.cgb_ideal_relu(self.handle, _ptr(self._u64(a0)), _ptr(self._u64(a1)), _ptr(out), a0.numel()))
        return out

    def ideal_relu_grad(self, g0, g1, z0, z1):
        out = self.torch.empty_like(g0)
        self.check(self.lib.cgb_ideal_relu_grad(self.handle, _ptr(self._u64(g0)), _ptr(self._u64(g1)), _ptr(self._u64(z0)),
                                                _ptr(self._u64(z1)), _ptr(out), g0.numel()))
        return out

    def ideal_softmax(self, z0, z1, labels, train_rows, f=SCALER_BITS):
        n, Cc = z0.shape
        assert labels.dtype == self.torch.int32 and labels.is_cuda
        P, pmy = self.torch.empty_like(z0), self.torch.empty_like(z0)
        self.check(self.lib.cgb_ideal_softmax(self.handle, _ptr(self._u64(z0)), _ptr(self._u64(z1)), _ptr(labels), n, Cc,
                                              train_rows, f, _ptr(P), _ptr(pmy)))
        return P, pmy

    def set_prg_stream_bias(self, bias_tensor):
        """bias_tensor: 1-element cuda int64 tensor (kept alive by the caller) or None."""
        self.check(self.lib.cgb_ctx_set_prg_stream_bias(self.handle, _ptr(bias_tensor)))

    # -- (4) PRG -----------------------------------------------------------------------------------------------
    def prg_fill(self, key, stream, word_offset, n_words, out=None):
        if out is None:
            out = self.empty(n_words)
        self.check(self.lib.cgb_prg_fill(self.handle, _lib.key_array(key), stream, word_offset, _ptr(out), n_words))
        return out

    def prediction_metrics(self, s0, s1, labels, train_rows, val_rows, f=SCALER_BITS):
        """(loss sum, hits full, hits train, hits test) from the two shares of the n x C probabilities and int32 labels."""
        n, Cc = s0.shape
        nb = C.c_uint32()
        self.check(self.lib.cgb_prediction_metrics(self.handle, None, None, None, n, Cc, train_rows, val_rows, f, None, C.byref(nb)))
        out = self.torch.zeros(max(nb.value, 1) * 4, dtype=self.torch.float64, device=self._dev())
        self.check(self.lib.cgb_prediction_metrics(self.handle, _ptr(self._u64(s0)), _ptr(self._u64(s1)), _ptr(labels), n, Cc,
                                                   train_rows, val_rows, f, _ptr(out), C.byref(nb)))
        rec = out.cpu().numpy().reshape(-1, 4)
        tot = [0.0, 0.0, 0.0, 0.0]
        for r in rec:  # in block order, like the engine's host side
            for k in range(4):
                tot[k] += float(r[k])
        return tuple(tot)

    def prg_fill_multi(self, key, segs):
        """segs: [(out, stream_a)] or [(out, stream_a, stream_b, out_b or None)]: out = PRG(a) [+ PRG(b)], out_b = PRG(b)."""
        m = 2**64 - 1
        arr = (_lib.PrgSeg * max(len(segs), 1))()
        for i, sg in enumerate(segs):
            out, a = sg[0], sg[1]
            arr[i].out, arr[i].n_words, arr[i].stream_a = self._u64(out).data_ptr(), out.numel(), int(a) & m
            if len(sg) > 2:
                arr[i].stream_b, arr[i].has_b = int(sg[2]) & m, 1
                arr[i].out_b = self._u64(sg[3]).data_ptr() if sg[3] is not None else None
        self.check(self.lib.cgb_prg_fill_multi(self.handle, _lib.key_array(key), arr, len(segs)))

    def rowmul_sub(self, a, b, c, out=None):
        rows, D = a.shape
        out = self.torch.empty_like(a) if out is None else out
        self.check(self.lib.cgb_rowmul_sub(self.handle, _ptr(self._u64(a)), _ptr(self._u64(b)), _ptr(self._u64(c)), _ptr(out),
                                           rows, D))
        return out

    def prg_sum(self, key, streams, tensors, n_words=None, out=None):
        """out = sum(tensors) + sum_k PRG(key, streams[k]) (word offset 0)."""
        n = n_words if n_words is not None else tensors[0].numel()
        out = self.empty(n) if out is None else out
        ns, ni = len(streams), len(tensors)
        st = (C.c_uint64 * max(ns, 1))(*[int(x) & (2**64 - 1) for x in streams])
        ptrs = (C.c_void_p * max(ni, 1))(*[C.c_void_p(self._u64(t).data_ptr()) for t in tensors])
        self.check(self.lib.cgb_prg_sum(self.handle, _lib.key_array(key), st, ns, ptrs, ni, _ptr(out), n))
        return out

    def prg_mask_sub(self, key, stream, word_offset, x, out=None):
        out = self.torch.empty_like(x) if out is None else out
        self.check(self.lib.cgb_prg_mask_sub(self.handle, _lib.key_array(key), stream, word_offset,
                                             _ptr(self._u64(x)), _ptr(out), x.numel()))
        return out
