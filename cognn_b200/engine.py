"""ctypes binding of libcognn_b200_host.so (include/cognn_b200_engine.h): the C++ host engine.  Plumbing only."""
import ctypes as C
import os

import numpy as np

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libcognn_b200_host.so")
_host = None


class CgeConfig(C.Structure):
    _fields_ = [("num_layers", C.c_int32), ("num_labels", C.c_int32), ("input_dim", C.c_int32),
                ("hidden_dim", C.c_int32), ("num_samples", C.c_int32), ("num_edges", C.c_int32),
                ("learning_rate", C.c_double), ("train_ratio", C.c_double), ("val_ratio", C.c_double),
                ("test_ratio", C.c_double), ("scaler_bits", C.c_int32), ("key", C.c_uint32 * 8),
                ("record_messages", C.c_int32), ("verbose", C.c_int32)]


ENGINE_SYMBOLS = ["cge_last_error", "cge_nccl_unique_id", "cge_create_loopback", "cge_create_nccl", "cge_destroy",
                  "cge_add_party", "cge_setup", "cge_run", "cge_download", "cge_message_count", "cge_message_info",
                  "cge_message_data", "cge_words_sent", "cge_rounds", "cge_launch_count", "cge_seconds_online", "cge_seconds_online_gpu", "cge_seconds_offline", "cge_seconds_residual_host",
                  "cge_metrics_count", "cge_metrics_get", "cge_build_party_graph", "cge_graph_replays", "cge_plane"]


def load_host():
    global _host
    if _host is not None:
        return _host
    _lib.load()
    if not os.path.exists(HOST_LIB_PATH):
        raise _lib.CgbError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
    h = C.CDLL(HOST_LIB_PATH)
    h.cge_last_error.restype = C.c_char_p
    h.cge_last_error.argtypes = [C.c_void_p]
    h.cge_download.restype = C.c_int64
    h.cge_download.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32),
                               C.POINTER(C.c_uint32)]
    for n in ("cge_message_count", "cge_words_sent", "cge_rounds", "cge_launch_count", "cge_metrics_count", "cge_graph_replays"):
        getattr(h, n).restype = C.c_uint64
        getattr(h, n).argtypes = [C.c_void_p]
    for n in ("cge_seconds_online", "cge_seconds_online_gpu", "cge_seconds_offline", "cge_seconds_residual_host"):
        getattr(h, n).restype = C.c_double
        getattr(h, n).argtypes = [C.c_void_p]
    h.cge_create_loopback.argtypes = [C.c_int, C.c_void_p, C.c_int, C.POINTER(CgeConfig), C.POINTER(C.c_void_p)]
    h.cge_create_nccl.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(CgeConfig), C.POINTER(C.c_void_p)]
    h.cge_destroy.argtypes = [C.c_void_p]
    h.cge_add_party.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    h.cge_setup.argtypes = [C.c_void_p]
    h.cge_run.argtypes = [C.c_void_p, C.c_uint64]
    h.cge_message_info.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    h.cge_message_data.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    h.cge_metrics_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int)] + [C.POINTER(C.c_double)] * 4
    h.cge_build_party_graph.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int] + [C.c_void_p] * 6 + \
        [C.POINTER(C.c_uint64)] * 3
    h.cge_nccl_unique_id.argtypes = [C.c_void_p]
    h.cge_plane.restype = C.c_char_p
    h.cge_plane.argtypes = [C.c_void_p]
    _host = h
    return h


def make_config(cfg, f=16, key=None, record=False, verbose=False):
    c = CgeConfig()
    c.num_layers = 2
    c.num_labels, c.input_dim, c.hidden_dim = cfg["num_labels"], cfg["input_dim"], cfg["hidden_dim"]
    c.num_samples, c.num_edges = cfg.get("num_samples", 0), cfg.get("num_edges", 0)
    c.learning_rate, c.train_ratio, c.val_ratio = cfg["learning_rate"], cfg["train_ratio"], cfg["val_ratio"]
    c.test_ratio = cfg.get("test_ratio", 1.0 - cfg["train_ratio"] - cfg["val_ratio"])
    c.scaler_bits = f
    for i, k in enumerate(key or [45, 0, 0, 0, 0, 0, 0, 0]):
        c.key[i] = k
    c.record_messages, c.verbose = int(record), int(verbose)
    return c


def build_party_graph(edges, tid, T, me):
    """Host-only index-vector builder of the C++ engine (ssk.h:295-534 with -r 1)."""
    h = load_host()
    edges = np.ascontiguousarray(edges, dtype=np.int64)
    tid = np.ascontiguousarray(tid, dtype=np.int64)
    nl, nr, nc = C.c_uint64(), C.c_uint64(), C.c_uint64()
    args = [edges.ctypes.data, edges.shape[0], tid.ctypes.data, tid.size, T, me]
    rc = h.cge_build_party_graph(*args, None, None, None, None, None, None, C.byref(nl), C.byref(nr), C.byref(nc))
    if rc != 0:
        raise _lib.CgbError(h.cge_last_error(None).decode())
    vids = np.zeros(nl.value, dtype=np.uint64)
    in_raw, in_deg = np.zeros(nl.value, dtype=np.uint64), np.zeros(nl.value, dtype=np.uint64)
    offsets = np.zeros(T + 1, dtype=np.uint32)
    rowptr, col = np.zeros(nr.value + 1, dtype=np.uint32), np.zeros(nc.value, dtype=np.uint32)
    rc = h.cge_build_party_graph(*args, vids.ctypes.data, in_raw.ctypes.data, in_deg.ctypes.data, offsets.ctypes.data,
                                 rowptr.ctypes.data, col.ctypes.data, C.byref(nl), C.byref(nr), C.byref(nc))
    if rc != 0:
        raise _lib.CgbError(h.cge_last_error(None).decode())
    return {"vids": vids, "in_deg_raw": in_raw, "in_deg": in_deg, "offsets": offsets, "rowptr": rowptr, "col": col}


class Engine:
    """All parties in one process (loopback) or one party of an NCCL group.  See include/cognn_b200_engine.h."""

    def __init__(self, T, cfg, device=0, f=16, key=None, record=False, verbose=False, rank=None, nccl_uid=None, stream=None):
        self.h = load_host()
        self.T, self.cfg = T, cfg
        c = make_config(cfg, f, key, record, verbose)
        out = C.c_void_p()
        if rank is None:
            rc = self.h.cge_create_loopback(device, C.c_void_p(stream) if stream else None, T, C.byref(c), C.byref(out))
            self.local = list(range(T))
        else:
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_uid))
            rc = self.h.cge_create_nccl(device, C.c_void_p(stream) if stream else None, rank, T, buf, C.byref(c), C.byref(out))
            self.local = [rank]
        if rc != 0:
            raise _lib.CgbError(self.h.cge_last_error(None).decode())
        self.e = out

    def _ck(self, rc):
        if rc != 0:
            raise _lib.CgbError(self.h.cge_last_error(self.e).decode())

    def load(self, edges, tid, feats, labels):
        edges = np.ascontiguousarray(edges, dtype=np.int64)
        tid = np.ascontiguousarray(tid, dtype=np.int64)
        feats = np.ascontiguousarray(feats, dtype=np.float64)
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        for p in self.local:
            self._ck(self.h.cge_add_party(self.e, p, edges.ctypes.data, edges.shape[0], tid.ctypes.data, tid.size,
                                          feats.ctypes.data, labels.ctypes.data))
        self._ck(self.h.cge_setup(self.e))

    def run(self, n_iters):
        self._ck(self.h.cge_run(self.e, n_iters))

    def download(self, owner, role, name):
        r, c = C.c_uint32(), C.c_uint32()
        n = self.h.cge_download(self.e, owner, role, name.encode(), None, 0, C.byref(r), C.byref(c))
        if n < 0:
            self._ck(-1)
        out = np.zeros(n, dtype=np.uint64)
        n = self.h.cge_download(self.e, owner, role, name.encode(), out.ctypes.data, out.size, C.byref(r), C.byref(c))
        if n < 0:
            self._ck(-1)
        return out.reshape(r.value, c.value)

    def messages(self):
        out = []
        for i in range(self.h.cge_message_count(self.e)):
            it, src, dst, nw = C.c_uint64(), C.c_int(), C.c_int(), C.c_uint64()
            tag = C.create_string_buffer(64)
            self._ck(self.h.cge_message_info(self.e, i, C.byref(it), C.byref(src), C.byref(dst), tag, 64, C.byref(nw)))
            data = np.zeros(nw.value, dtype=np.uint64)
            self._ck(self.h.cge_message_data(self.e, i, data.ctypes.data, data.size))
            out.append((it.value, src.value, dst.value, tag.value.decode(), data))
        return out

    def metrics(self):
        out = []
        for i in range(self.h.cge_metrics_count(self.e)):
            it, party = C.c_uint64(), C.c_int()
            v = [C.c_double() for _ in range(4)]
            self._ck(self.h.cge_metrics_get(self.e, i, C.byref(it), C.byref(party), *[C.byref(x) for x in v]))
            out.append({"iter": it.value, "party": party.value, "loss": v[0].value, "acc_full": v[1].value,
                        "acc_train": v[2].value, "acc_test": v[3].value})
        return out

    @property
    def words_sent(self):
        return int(self.h.cge_words_sent(self.e))

    @property
    def plane(self):
        """The message plane the rounds run on (decided in setup: peer memory when every rank could map its peers)."""
        return self.h.cge_plane(self.e).decode()

    @property
    def rounds(self):
        return int(self.h.cge_rounds(self.e))

    @property
    def seconds_online(self):
        return float(self.h.cge_seconds_online(self.e))

    @property
    def seconds_online_gpu(self):
        return float(self.h.cge_seconds_online_gpu(self.e))

    @property
    def seconds_residual_host(self):
        return float(self.h.cge_seconds_residual_host(self.e))

    @property
    def seconds_offline(self):
        return float(self.h.cge_seconds_offline(self.e))

    @property
    def launches(self):
        return int(self.h.cge_launch_count(self.e))

    @property
    def graph_replays(self):
        return int(self.h.cge_graph_replays(self.e))

    def close(self):
        if self.e is not None:
            self.h.cge_destroy(self.e)
            self.e = None
