"""ctypes binding of libcognn_b200.so (the C ABI declared in include/cognn_b200.h).

Only plumbing lives here: tests and bench.py reach the CUDA kernels through exactly the entry points a CoGNN
build would bind.  The library is required -- there is no Python or CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# COGNN_B200_LIB: load another build of the same C ABI (A/B runs of compile-time variants, e.g. -DCGB_CHUNK_EDGES=128u)
LIB_PATH = os.environ.get("COGNN_B200_LIB") or os.path.join(_HERE, "libcognn_b200.so")

u64p = C.c_void_p  # device / host pointers travel as integers
u32p = C.c_void_p
ctx_p = C.c_void_p
csr_p = C.c_void_p


class CgbError(RuntimeError):
    pass


# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header is exported
SIGNATURES = {
    "cgb_version": (C.c_char_p, []),
    "cgb_device_count": (C.c_int, []),
    "cgb_ctx_create": (C.c_int, [C.c_int, C.POINTER(ctx_p)]),
    "cgb_ctx_create_on_stream": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(ctx_p)]),
    "cgb_ctx_destroy": (C.c_int, [ctx_p]),
    "cgb_ctx_sync": (C.c_int, [ctx_p]),
    "cgb_ctx_stream": (C.c_void_p, [ctx_p]),
    "cgb_last_error": (C.c_char_p, [ctx_p]),
    "cgb_ctx_launch_count": (C.c_uint64, [ctx_p]),
    "cgb_ctx_set_prg_stream_bias": (C.c_int, [ctx_p, C.c_void_p]),
    "cgb_malloc": (C.c_int, [ctx_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "cgb_free": (C.c_int, [ctx_p, C.c_void_p]),
    "cgb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "cgb_host_free": (C.c_int, [C.c_void_p]),
    "cgb_memset": (C.c_int, [ctx_p, C.c_void_p, C.c_int, C.c_size_t]),
    "cgb_h2d": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "cgb_d2h": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "cgb_d2d": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "cgb_csr_create": (C.c_int, [ctx_p, u32p, u32p, C.c_uint32, C.c_uint64, C.c_uint32, C.POINTER(csr_p)]),
    "cgb_csr_create_device": (C.c_int, [ctx_p, u32p, u32p, C.c_uint32, C.c_uint64, C.c_uint32, C.POINTER(csr_p)]),
    "cgb_csr_destroy": (C.c_int, [ctx_p, csr_p]),
    "cgb_csr_num_edges": (C.c_uint64, [csr_p]),
    "cgb_csr_num_rows": (C.c_uint32, [csr_p]),
    "cgb_csr_rowptr": (C.c_void_p, [csr_p]),
    "cgb_csr_col": (C.c_void_p, [csr_p]),
    "cgb_gather_sum": (C.c_int, [ctx_p, csr_p, u64p, u64p, u64p, C.c_uint32]),
    "cgb_gather_sum_blocks": (C.c_int, [ctx_p, csr_p, u64p, u64p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "cgb_gather_sum_compact": (C.c_int, [ctx_p, csr_p, u64p, u64p, u64p, C.c_uint32]),
    "cgb_gather_sum_signal": (C.c_int, [ctx_p, csr_p, u64p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint32]),
    "cgb_csr_num_nonempty_rows": (C.c_uint32, [csr_p]),
    "cgb_csr_nonempty_rows": (C.c_void_p, [csr_p]),
    "cgb_scatter_add_rows": (C.c_int, [ctx_p, u32p, C.c_uint64, u64p, u64p, C.c_uint32, C.c_int, C.c_uint32]),
    "cgb_flag_signal": (C.c_int, [ctx_p, C.c_void_p, C.c_uint32]),
    "cgb_flag_wait": (C.c_int, [ctx_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]),
    "cgb_ctx_last_kernel": (C.c_char_p, [ctx_p]),
    "cgb_ctx_set_matmul_impl": (C.c_int, [ctx_p, C.c_int]),
    "cgb_probe_imad_peak": (C.c_int, [ctx_p, C.POINTER(C.c_double)]),
    "cgb_probe_tensor_i8_peak": (C.c_int, [ctx_p, C.POINTER(C.c_double)]),
    "cgb_ipc_export": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p]),
    "cgb_peer_round": (C.c_int, [ctx_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32,
                                 C.c_void_p]),
    "cgb_ipc_open": (C.c_int, [ctx_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cgb_ipc_close": (C.c_int, [ctx_p, C.c_void_p]),
    "cgb_peer_copy": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32]),
    "cgb_party_graph_build": (C.c_int, [ctx_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "cgb_party_graph_build_host": (C.c_int, [ctx_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                             C.POINTER(C.c_void_p)]),
    "cgb_party_graph_destroy": (C.c_int, [ctx_p, C.c_void_p]),
    "cgb_party_graph_num_local": (C.c_uint32, [C.c_void_p]),
    "cgb_party_graph_num_rows": (C.c_uint32, [C.c_void_p]),
    "cgb_party_graph_num_out_edges": (C.c_uint64, [C.c_void_p]),
    "cgb_party_graph_offsets": (C.c_void_p, [C.c_void_p]),
    "cgb_party_graph_vids": (C.c_void_p, [C.c_void_p]),
    "cgb_party_graph_in_deg_raw": (C.c_void_p, [C.c_void_p]),
    "cgb_party_graph_in_deg": (C.c_void_p, [C.c_void_p]),
    "cgb_party_graph_is_border": (C.c_void_p, [C.c_void_p]),
    "cgb_party_graph_rowptr": (C.c_void_p, [C.c_void_p]),
    "cgb_party_graph_col": (C.c_void_p, [C.c_void_p]),
    "cgb_party_graph_csr": (C.c_int, [ctx_p, C.c_void_p, C.POINTER(csr_p)]),
    "cgb_expand_rows": (C.c_int, [ctx_p, u32p, C.c_uint64, u64p, u64p, u64p, C.c_uint32]),
    "cgb_segsum": (C.c_int, [ctx_p, u32p, C.c_uint32, C.c_uint64, u64p, u64p, C.c_uint32, C.c_int]),
    "cgb_matmul": (C.c_int, [ctx_p, u64p, u64p, u64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int]),
    "cgb_beaver_matmul_finish": (C.c_int, [ctx_p, u64p, u64p, u64p, u64p, u64p, u64p, C.c_uint32, C.c_uint32,
                                           C.c_uint32, C.c_int, C.c_int]),
    "cgb_add": (C.c_int, [ctx_p, u64p, u64p, u64p, C.c_uint64]),
    "cgb_sub": (C.c_int, [ctx_p, u64p, u64p, u64p, C.c_uint64]),
    "cgb_sum_n": (C.c_int, [ctx_p, C.c_void_p, C.c_uint32, u64p, C.c_uint64]),
    "cgb_trunc": (C.c_int, [ctx_p, u64p, u64p, C.c_uint64, C.c_int, C.c_int]),
    "cgb_scale_public": (C.c_int, [ctx_p, u64p, C.c_uint64, u64p, C.c_uint64, C.c_int, C.c_int]),
    "cgb_apply_gradient": (C.c_int, [ctx_p, u64p, u64p, C.c_uint64, u64p, C.c_uint64, C.c_int, C.c_int]),
    "cgb_rowmul_beaver_finish": (C.c_int, [ctx_p, u64p, u64p, u64p, u64p, u64p, u64p, C.c_uint64, C.c_uint32,
                                           C.c_int, C.c_int]),
    "cgb_rowmul_beaver_finish_open": (C.c_int, [ctx_p, u64p, u64p, u64p, u64p, u64p, u64p, C.c_uint64, C.c_uint32,
                                                C.c_int, C.c_int]),
    "cgb_sub_pair": (C.c_int, [ctx_p, u64p, u64p, C.c_uint64, u64p, u64p, C.c_uint64, u64p]),
    "cgb_beaver_matmul_finish_open": (C.c_int, [ctx_p, u64p, u64p, u64p, u64p, u64p, u64p, C.c_uint32, C.c_uint32,
                                                C.c_uint32, C.c_int, C.c_int]),
    "cgb_prg_sum": (C.c_int, [ctx_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p, C.c_uint32, u64p,
                              C.c_uint64]),
    "cgb_scale_apply_gradient": (C.c_int, [ctx_p, u64p, u64p, C.c_uint64, C.c_uint64, u64p, u64p, C.c_uint64, C.c_int, C.c_int]),
    "cgb_avg_public": (C.c_int, [ctx_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_int]),
    "cgb_ideal_relu_reshare": (C.c_int, [ctx_p, C.POINTER(C.c_uint32), C.c_uint64, u64p, u64p, u64p, u64p, u64p, C.c_uint64]),
    "cgb_prediction_metrics": (C.c_int, [ctx_p, u64p, u64p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                         C.c_void_p, C.POINTER(C.c_uint32)]),
    "cgb_prg_fill_multi": (C.c_int, [ctx_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32]),
    "cgb_rowmul_sub": (C.c_int, [ctx_p, u64p, u64p, u64p, u64p, C.c_uint64, C.c_uint32]),
    "cgb_copy_segments": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32]),
    "cgb_cond_add": (C.c_int, [ctx_p, u64p, u64p, C.c_void_p, u64p, C.c_uint64, C.c_uint32]),
    "cgb_transpose": (C.c_int, [ctx_p, u64p, u64p, C.c_uint32, C.c_uint32]),
    "cgb_encode": (C.c_int, [ctx_p, C.c_void_p, u64p, C.c_uint64, C.c_int]),
    "cgb_decode": (C.c_int, [ctx_p, u64p, C.c_void_p, C.c_uint64, C.c_int]),
    "cgb_share_split": (C.c_int, [ctx_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint32), C.c_uint64,
                                  C.c_uint64, u64p, u64p]),
    "cgb_open_decode": (C.c_int, [ctx_p, u64p, u64p, C.c_void_p, C.c_uint64, C.c_int]),
    "cgb_prg_fill": (C.c_int, [ctx_p, C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64, u64p, C.c_uint64]),
    "cgb_prg_mask_sub": (C.c_int, [ctx_p, C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64, u64p, u64p, C.c_uint64]),
    "cgb_ideal_relu": (C.c_int, [ctx_p, u64p, u64p, u64p, C.c_uint64]),
    "cgb_ideal_relu_grad": (C.c_int, [ctx_p, u64p, u64p, u64p, u64p, u64p, C.c_uint64]),
    "cgb_ideal_softmax": (C.c_int, [ctx_p, u64p, u64p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, u64p, u64p]),
    "cgb_host_gather_sum": (C.c_int, [ctx_p, csr_p, u64p, u64p, u64p, C.c_uint32]),
    "cgb_host_gather_sum_async": (C.c_int, [ctx_p, csr_p, u64p, u64p, u64p, C.c_uint32]),
    "cgb_host_sync": (C.c_int, [ctx_p]),
}

_lib = None


def load():
    """Loads the shared library; raises loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CgbError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(cognn_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PrgSeg(C.Structure):
    """cgb_prg_seg (include/cognn_b200.h)"""
    _fields_ = [("out", C.c_void_p), ("out_b", C.c_void_p), ("n_words", C.c_uint64), ("stream_a", C.c_uint64),
                ("stream_b", C.c_uint64), ("has_b", C.c_uint32)]


class XSeg(C.Structure):
    """cgb_xseg (include/cognn_b200.h)"""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n_words", C.c_uint64), ("link", C.c_uint32)]


class XLink(C.Structure):
    """cgb_xlink (include/cognn_b200.h)"""
    _fields_ = [("seq", C.c_void_p), ("done", C.c_void_p), ("wait_flag", C.c_void_p), ("signal_flag", C.c_void_p),
                ("recv", C.c_uint32)]


def key_array(key):
    assert len(key) == 8
    return (C.c_uint32 * 8)(*[int(k) & 0xFFFFFFFF for k in key])
