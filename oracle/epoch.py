"""CPU ORACLE (test infrastructure) for one secure-GCN epoch / inference of the CoGNN-Opt operators, all T parties in
one address space.  Restates the share-local dataflow of

  /root/reference/include/ss_vertex_centric_algo_kernel.h:680-1189   (onIteration / runAlgoKernelServer)
  /root/reference/algo_kernels/vertex_centric/optimize-gcn/gcn.h:198-887   (PreScatter/Scatter/Gather/Apply, weights)

on flat u64 arrays with the protocol frozen in DESIGN.md ("Protocol"): Beaver triples for every multiplication,
OM-style masked fused gather for Scatter/Gather, SecureML local truncation, dealer randomness from the ChaCha20 PRG.
The 2PC-RESIDUAL steps (ReLU, softmax/p-y, ReLU') are an IDEAL-FUNCTIONALITY stand-in on the host: the helper's share
is sent to the owner, the function is evaluated in the clear and re-shared.  That stand-in is NOT secure and exists
only so an epoch can run end to end; the reference keeps these steps on its SCI/OT backend.

PARITY UNPINNED by the reference (no tests / vectors there).  Pinned here: reconstructed values against a float64
GCN of the same dataflow (PlainGCN below), index vectors against SURVEY.md 3.6, glibc rand() weights.
Used by tests/ to check the C++/CUDA engine bit for bit: final shares of every party and every message.
"""
import ctypes
import math

import numpy as np

from . import graph_index as gi
from . import pyoracle as po

DEFAULT_KEY = [45, 0, 0, 0, 0, 0, 0, 0]
U64 = np.uint64

# PRG stream kinds (DESIGN.md "Randomness")
K_FEAT, K_WEIGHT, K_OM_R, K_OM_S = 1, 2, 3, 4
K_MM_U0, K_MM_U1, K_MM_V0, K_MM_V1, K_MM_Z0 = 5, 6, 7, 8, 9
K_RM_A0, K_RM_A1, K_RM_B0, K_RM_B1, K_RM_C0 = 10, 11, 12, 13, 14
K_RESHARE = 15


def stream_id(kind, it, owner, sub=0):
    assert owner < 256 and sub < 256 and it < (1 << 32)
    return (kind << 48) | (it << 16) | (owner << 8) | sub


def glibc_init_weight(d0, d1):
    """optimize-gcn/gcn.h:838-852: std::srand(42) per call, (double) rand() / RAND_MAX * 2 * limit - limit."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(42)
    limit = math.sqrt(6.0 / (d0 + d1))
    W = np.zeros((d0, d1), dtype=np.float64)
    for i in range(d0):
        for j in range(d1):
            W[i, j] = float(libc.rand()) / 2147483647.0 * 2 * limit - limit
    return W


def norm_vector(deg, f):
    """gcn.h:219-221, 472-474, 537-539: deg == 0 ? 0 : encodeDoubleAsFixedPoint(pow(deg + 1, -0.5))."""
    out = np.zeros(len(deg), dtype=U64)
    for i, d in enumerate(np.asarray(deg).tolist()):
        out[i] = 0 if d == 0 else po.lib().orc_encode_fixed(math.pow(float(d) + 1.0, -0.5), f)
    return out


def fused_csr(iv, all_ivs):
    """One CSR over every out-edge of a party: rows = destination vertices of party 0, then party 1, ... (each in its
    owner's localVertexPos order), columns = local source rows.  Dummy entries (ssk.h:412-418) are dropped: they only
    exist so the helper cannot see a zero in-degree, and here the helper never sees the edge layout at all."""
    T = iv["T"]
    rowptrs, cols, off = [np.zeros(1, dtype=np.int64)], [], 0
    for t in range(T):
        rp, col = gi.csr_from_pos(iv["updateSrcVertexPos"][t], iv["updateDstVertexPos"][t], iv["localVertexPos"],
                                  all_ivs[t]["localVertexPos"], drop=iv["isUpdateSrcVertexDummy"][t])
        rowptrs.append(rp[1:].astype(np.int64) + off)
        off += int(rp[-1])
        cols.append(col)
    return np.concatenate(rowptrs).astype(np.uint32), (np.concatenate(cols) if cols else np.zeros(0)).astype(np.uint32)


class EpochOracle:
    def __init__(self, edges, tid_of, T, feats, labels, cfg, f=16, key=None):
        self.T, self.f, self.key, self.cfg = T, f, list(key or DEFAULT_KEY), dict(cfg)
        self.F, self.H, self.C = cfg["input_dim"], cfg["hidden_dim"], cfg["num_labels"]
        self.tiles, self.ivs = gi.build_all(edges, tid_of, T, no_dummy_edge=True)
        self.msgs = []  # (iteration, src, dst, tag, flat u64 copy) in program order
        self.log = []
        feats = np.asarray(feats, dtype=np.float64)
        labels = np.asarray(labels, dtype=np.int64)
        W_plain = [glibc_init_weight(self.F, self.H), glibc_init_weight(self.H, self.C)]
        self.W_plain_init = W_plain
        self.own, self.hlp = [], []  # share 0 state at party p; share 1 state of owner p (lives on party q(p))
        self.n, self.norm, self.csr, self.labels, self.offsets = [], [], [], [], []
        for p in range(T):
            iv, tile = self.ivs[p], self.tiles[p]
            vids = iv["localVertexPos"].astype(np.int64)
            self.n.append(len(vids))
        self.offsets = np.concatenate([[0], np.cumsum(self.n)]).astype(np.int64)
        for p in range(T):
            iv, tile = self.ivs[p], self.tiles[p]
            vids = iv["localVertexPos"].astype(np.int64)
            # gcn.h:819-835, 857-862: raw * (inDeg + 1)^-1/2 with the in-degree BEFORE the dummy increment (ssk.h:177 < 190)
            scale = np.array([math.pow(float(d) + 1.0, -0.5) for d in tile.in_deg[vids].tolist()])
            x = feats[vids] * scale[:, None]
            x0, x1 = po.share_split(x, f, self.key, stream_id(K_FEAT, 0, p))
            own = {"X": x0, "X_backup": x0.copy(), "W": [], "h_t": [None, None], "z": [None, None], "g": None}
            hlp = {"X": x1, "X_backup": x1.copy(), "W": [], "h_t": [None, None], "z": [None, None], "g": None}
            for l in range(2):
                w0, w1 = po.share_split(W_plain[l], f, self.key, stream_id(K_WEIGHT, 0, p, l))
                own["W"].append(w0)
                hlp["W"].append(w1)
            self.own.append(own)
            self.hlp.append(hlp)
            self.norm.append(norm_vector(iv["localVertexInDeg"], f))  # PreScatter is handed inDeg too (ssk.h:739)
            self.csr.append(fused_csr(iv, self.ivs))
            self.labels.append(labels[vids])
        self.lr = po.lib().orc_encode_fixed(cfg["learning_rate"], f)

    # ---- plumbing ---------------------------------------------------------------------------------------------
    def q(self, p):
        return (p + 1) % self.T

    def prg(self, kind, it, owner, sub, shape):
        n = int(np.prod(shape))
        return po.prg_fill(self.key, stream_id(kind, it, owner, sub), 0, n).reshape(shape)

    def send(self, it, src, dst, tag, *arrays):
        flat = np.concatenate([np.ascontiguousarray(a, dtype=U64).ravel() for a in arrays])
        self.msgs.append((it, src, dst, tag, flat.copy()))

    # ---- two-party building blocks (owner p = share 0, helper q(p) = share 1) -----------------------------------
    def beaver_matmul(self, it, p, sub, A0, A1, B0, B1):
        """sci::twoPartyGCNMatMul (gcn.h:233,665,671,710): C = trunc(A * B)."""
        q, f = self.q(p), self.f
        (M, K), N = A0.shape, B0.shape[1]
        U0, U1 = self.prg(K_MM_U0, it, p, sub, (M, K)), self.prg(K_MM_U1, it, p, sub, (M, K))
        V0, V1 = self.prg(K_MM_V0, it, p, sub, (K, N)), self.prg(K_MM_V1, it, p, sub, (K, N))
        Z0 = self.prg(K_MM_Z0, it, p, sub, (M, N))
        Z1 = po.sub(po.matmul(po.add(U0, U1), po.add(V0, V1)), Z0)  # dealer (offline)
        E0, F0 = po.sub(A0, U0), po.sub(B0, V0)
        E1, F1 = po.sub(A1, U1), po.sub(B1, V1)
        self.send(it, p, q, f"mm{sub}.o{p}", E0, F0)
        self.send(it, q, p, f"mm{sub}.o{p}", E1, F1)
        E, F = po.add(E0, E1), po.add(F0, F1)
        return (po.beaver_matmul_finish(E, F, U0, V0, Z0, 0, f), po.beaver_matmul_finish(E, F, U1, V1, Z1, 1, f))

    def rowmul(self, it, p, sub, x0, x1, s):
        """sci::twoPartyGCNVectorScale (gcn.h:247,476): y = trunc(x * s[row]), s private to the owner."""
        q, f = self.q(p), self.f
        rows, D = x0.shape
        a0, a1 = self.prg(K_RM_A0, it, p, sub, (rows, D)), self.prg(K_RM_A1, it, p, sub, (rows, D))
        b0, b1 = self.prg(K_RM_B0, it, p, sub, (rows,)), self.prg(K_RM_B1, it, p, sub, (rows,))
        c0 = self.prg(K_RM_C0, it, p, sub, (rows, D))
        c1 = (a0 + a1) * (b0 + b1)[:, None] - c0  # dealer (offline)
        e0, f0 = po.sub(x0, a0), po.sub(s, b0)
        e1, f1 = po.sub(x1, a1), po.sub(np.zeros(rows, dtype=U64), b1)
        self.send(it, p, q, f"rm{sub}.o{p}", e0, f0)
        self.send(it, q, p, f"rm{sub}.o{p}", e1, f1)
        e, fv = po.add(e0, e1), po.add(f0, f1)
        return (po.rowmul_beaver_finish(e, fv, a0, b0, c0, 0, f), po.rowmul_beaver_finish(e, fv, a1, b1, c1, 1, f))

    def residual(self, it, p, sub, fn, ins0, ins1, n_out):
        """2PC-RESIDUAL stand-in (ideal functionality, NOT secure): helper sends its shares, owner evaluates in the
        clear on the host and re-shares; the helper's new share is a PRG stream both know from the dealer."""
        q = self.q(p)
        self.send(it, q, p, f"res{sub}.o{p}", *ins1)
        outs = fn(*[po.add(a, b) for a, b in zip(ins0, ins1)])
        assert len(outs) == n_out
        o0, o1 = [], []
        for k, v in enumerate(outs):
            s1 = self.prg(K_RESHARE, it, p, sub + k, v.shape)
            o1.append(s1)
            o0.append(po.sub(v, s1))
        return o0, o1

    # ---- Scatter / Gather of one GAS iteration (ssk.h:748-880 share-local composite) ------------------------------
    def gas(self, it, X0s, X1s):
        """V = X + sum over all in-edges (local and mirror) of X[src], for every owner.  X0s/X1s: per-owner shares."""
        T = self.T
        D = X0s[0].shape[1]
        Y = [None] * T
        for p in range(T):  # round 1: OM online message helper -> owner, then the fused gather at the owner
            q = self.q(p)
            r = self.prg(K_OM_R, it, p, 0, (self.n[p], D))
            m = po.sub(X1s[p], r)
            self.send(it, q, p, f"om.o{p}", m)
            S = np.concatenate([self.prg(K_OM_S, it, p, t, (self.n[t], D)) for t in range(T)])
            rowptr, col = self.csr[p]
            delta = po.sub(po.gather_sum_csr(rowptr, col, r), S)  # dealer (offline): A r - s
            Y[p] = po.gather_sum_csr(rowptr, col, po.add(X0s[p], m), delta)
        for p in range(T):  # round 2: mirror-update blocks to the primary helper of each destination owner
            for t in range(T):
                if t != p and self.q(t) != p:
                    self.send(it, p, self.q(t), f"upd.o{p}.t{t}", Y[p][self.offsets[t]:self.offsets[t + 1]])
        V0s, V1s = [], []
        for t in range(T):  # GatherComp additions (gcn.h:456-463): owner side and helper side
            lo, hi = self.offsets[t], self.offsets[t + 1]
            v0 = po.add(X0s[t], Y[t][lo:hi])                         # own local block
            v1 = po.add(X1s[t], self.prg(K_OM_S, it, t, t, (self.n[t], D)))
            for p in range(T):
                if p == t:
                    continue
                v0 = po.add(v0, self.prg(K_OM_S, it, p, t, (self.n[t], D)))  # owner holds the mask share s_{p->t}
                v1 = po.add(v1, Y[p][lo:hi])                                   # helper holds the masked sums
            V0s.append(v0)
            V1s.append(v1)
        return V0s, V1s

    # ---- 2PC-residual functions (host, float64 where the reference is float) -------------------------------------
    def f_relu(self, z):
        return [np.where(z.astype(np.int64) > 0, z, U64(0))]

    def make_f_softmax(self, p):
        f, C = self.f, self.C
        labels = self.labels[p]
        n = self.n[p]
        train = int(n * self.cfg["train_ratio"])  # gcn.h:560: (uint64_t)(vecSize * train_ratio)

        def fn(z):  # device stand-in of the engine: cgb_ideal_softmax (exp restated with IEEE + - * only, orc_det_exp)
            P, pmy = po.ideal_softmax(z, np.zeros_like(z), labels, train, f)
            return [P, pmy]

        return fn

    def f_relu_grad(self, g, z):
        return [np.where(z.astype(np.int64) > 0, g, U64(0))]

    # ---- weight averaging (gcn.h:747-802) -------------------------------------------------------------------------
    def weight_average(self, it, layer):
        T, f = self.T, self.f
        if T == 1:
            return
        # party 0 accumulates A0 = W0_0 + sum_{i>=1} W1_i (W1_i is held by party q(i)); party 1 accumulates
        # A1 = W0_1 + sum_{i>=2} W0_i + W1_0 (gcn.h:753-765 + 773-775)
        for i in range(2, T):
            self.send(it, i, 1, f"w{layer}.own{i}", self.own[i]["W"][layer])        # clientTaskComm.send(weightRef, 1)
            self.send(it, i, 0, f"w{layer}.hlp{i - 1}", self.hlp[i - 1]["W"][layer])    # serverTaskComm.send(coWeightRef, 0)
        A0 = self.own[0]["W"][layer].copy()
        for i in range(2, T):
            A0 = po.add(A0, self.hlp[i - 1]["W"][layer])
        A0 = po.add(A0, self.hlp[T - 1]["W"][layer])  # party 0's remoteWeight = share 1 of party T-1's replica
        A1 = self.own[1]["W"][layer].copy()
        for i in range(2, T):
            A1 = po.add(A1, self.own[i]["W"][layer])
        A1 = po.add(A1, self.hlp[0]["W"][layer])
        c = po.lib().orc_encode_fixed(1.0 / T, f)  # static_cast<uint64_t>(weightScaler * (1<<f)), gcn.h:763-764
        A0 = po.scale_public(A0, c, f, 0)
        A1 = po.scale_public(A1, c, f, 1)
        for i in range(2, T):
            self.send(it, 1, i, f"wavg{layer}.A1", A1)   # party i: weightRef <- party 1
            self.send(it, 0, i, f"wavg{layer}.A0", A0)   # party i: coWeightRef <- party 0
        # afterwards: party 0 holds (A0, A0), party 1 (A1, A1), party i >= 2 (local A1, remote A0)  (gcn.h:765-777)
        for p in range(T):
            self.own[p]["W"][layer] = (A0 if p == 0 else A1).copy()
            holder = self.q(p)  # hlp[p] lives on party q(p); that party's remote weight after averaging:
            self.hlp[p]["W"][layer] = (A1 if holder == 1 else A0).copy()

    # ---- iterations ------------------------------------------------------------------------------------------------
    def run(self, n_iters):
        """Runs the next n_iters GAS iterations (like the engine's run(): a second call continues where the first stopped)."""
        T, f = self.T, self.f
        start = getattr(self, "iters_done", 0)
        self.iters_done = start + n_iters
        for it in range(start, start + n_iters):
            ph = it % 6
            if ph == 0:  # ssk.h:695, 938: back to the first layer
                for p in range(T):
                    self.own[p]["X"] = self.own[p]["X_backup"].copy()
                    self.hlp[p]["X"] = self.hlp[p]["X_backup"].copy()
            if ph in (0, 1):
                self.forward(it, ph)
            elif ph == 2:
                self.backward_first(it)
            elif ph == 3:
                self.backward_gas(it, layer=1)
            elif ph == 4:
                self.backward_relu(it)
            else:
                self.backward_gas(it, layer=0)

    def forward(self, it, layer):
        T = self.T
        X0s, X1s = [], []
        for p in range(T):  # PreScatterComp (gcn.h:198-255)
            own, hlp = self.own[p], self.hlp[p]
            own["h_t"][layer] = po.transpose(own["X"])
            hlp["h_t"][layer] = po.transpose(hlp["X"])
            x0, x1 = self.beaver_matmul(it, p, 0, own["X"], hlp["X"], own["W"][layer], hlp["W"][layer])
            if layer != 0:  # gcn.h:243-254
                x0, x1 = self.rowmul(it, p, 0, x0, x1, self.norm[p])
            X0s.append(x0)
            X1s.append(x1)
        V0s, V1s = self.gas(it, X0s, X1s)
        for p in range(T):
            own, hlp = self.own[p], self.hlp[p]
            v0, v1 = self.rowmul(it, p, 1, V0s[p], V1s[p], self.norm[p])  # gcn.h:470-484 ((it+1) % 6 != 0 here)
            own["z"][layer], hlp["z"][layer] = v0, v1
            if layer == 0:  # gcn.h:546-558
                (h0,), (h1,) = self.residual(it, p, 0, self.f_relu, [v0], [v1], 1)
                own["X"], hlp["X"] = h0, h1
            else:  # gcn.h:559-642
                (P0, d0), (P1, d1) = self.residual(it, p, 0, self.make_f_softmax(p), [v0], [v1], 2)
                self.send(it, self.q(p), p, f"open_p.o{p}", P1)  # getPlainShareVecVec (gcn.h:604): the owner learns p
                prob = po.open_decode(P0, P1, self.f)
                self.log.append(self.metrics(it, p, prob))
                own["X"], hlp["X"] = d0, d1

    def metrics(self, it, p, prob):
        n = self.n[p]
        labels = self.labels[p]
        train = int(n * self.cfg["train_ratio"])
        val = int(n * self.cfg["val_ratio"])
        pp = np.where(prob == 0, 0.001, prob)  # gcn.h:615
        loss = float(-np.log(np.maximum(pp[np.arange(n), labels], 1e-30)).mean()) if n else 0.0
        pred = pp.argmax(axis=1)
        acc = lambda lo, hi: float((pred[lo:hi] == labels[lo:hi]).mean()) if hi > lo else 0.0
        return {"iter": it, "party": p, "loss": loss, "acc_full": acc(0, n), "acc_train": acc(0, train),
                "acc_test": acc(train + val, n)}

    def backward_first(self, it):
        """iter % 6 == 2, apply only (ssk.h:709-732; gcn.h:664-669): g = (p - y) * W1^T."""
        for p in range(self.T):
            own, hlp = self.own[p], self.hlp[p]
            g0, g1 = self.beaver_matmul(it, p, 0, own["X"], hlp["X"], po.transpose(own["W"][1]), po.transpose(hlp["W"][1]))
            own["g"], hlp["g"] = g0, g1

    def backward_gas(self, it, layer):
        """iter % 6 in (3, 5): GAS over the upstream gradient, then d = h_t * v, gradient step, FedAvg."""
        T, f = self.T, self.f
        X0s, X1s = [], []
        for p in range(T):  # PreScatterComp backward: scale only (gcn.h:247-254)
            x0, x1 = self.rowmul(it, p, 0, self.own[p]["X"], self.hlp[p]["X"], self.norm[p])
            X0s.append(x0)
            X1s.append(x1)
        V0s, V1s = self.gas(it, X0s, X1s)
        for p in range(T):
            own, hlp = self.own[p], self.hlp[p]
            v0, v1 = V0s[p], V1s[p]
            if (it + 1) % 6 != 0:  # gcn.h:470: no in-degree scaling on the last iteration of the epoch
                v0, v1 = self.rowmul(it, p, 1, v0, v1, self.norm[p])
            d0, d1 = self.beaver_matmul(it, p, 1, own["h_t"][layer], hlp["h_t"][layer], v0, v1)  # gcn.h:671, 710
            train = int(self.n[p] * self.cfg["train_ratio"])
            gs = po.lib().orc_encode_fixed(1.0 / train, f) if train else 0  # gcn.h:673-676
            d0, d1 = po.scale_public(d0, gs, f, 0), po.scale_public(d1, gs, f, 1)
            own["W"][layer] = po.apply_gradient(own["W"][layer], d0, self.lr, f, 0)  # gcn.h:678, 730
            hlp["W"][layer] = po.apply_gradient(hlp["W"][layer], d1, self.lr, f, 1)
            if layer == 1:
                own["X"], hlp["X"] = own["g"], hlp["g"]  # dstVec.swap(g) (gcn.h:684)
            else:
                own["X"], hlp["X"] = v0, v1  # first layer: g is empty in the reference; the value is never used again
        self.weight_average(it, layer)

    def backward_relu(self, it):
        """iter % 6 == 4, apply only (gcn.h:702-708): g . ReLU'(z0); first layer, so no further matmul."""
        for p in range(self.T):
            own, hlp = self.own[p], self.hlp[p]
            (x0,), (x1,) = self.residual(it, p, 0, self.f_relu_grad, [own["X"], own["z"][0]], [hlp["X"], hlp["z"][0]], 1)
            own["X"], hlp["X"] = x0, x1

    # ---- views for tests ---------------------------------------------------------------------------------------------
    def weights_plain(self, p, layer):
        return po.open_decode(self.own[p]["W"][layer], self.hlp[p]["W"][layer], self.f)

    def x_plain(self, p):
        return po.open_decode(self.own[p]["X"], self.hlp[p]["X"], self.f)


class PlainGCN:
    """float64 GCN with the same dataflow and quirks (independent check of the reconstructed oracle values)."""

    def __init__(self, oracle, feats, labels):
        o = self.o = oracle
        T = o.T
        self.vids = [o.ivs[p]["localVertexPos"].astype(np.int64) for p in range(T)]
        self.X = [np.asarray(feats, dtype=np.float64)[self.vids[p]] *
                  np.array([math.pow(float(d) + 1.0, -0.5) for d in o.tiles[p].in_deg[self.vids[p]].tolist()])[:, None]
                  for p in range(T)]
        self.nrm = [np.array([0.0 if d == 0 else math.pow(float(d) + 1.0, -0.5) for d in o.ivs[p]["localVertexInDeg"].tolist()])
                    for p in range(T)]
        self.W = [[w.copy() for w in o.W_plain_init] for _ in range(T)]
        self.labels = [np.asarray(labels)[self.vids[p]] for p in range(T)]

    def gas(self, Xs):
        o = self.o
        T = o.T
        V = [x.copy() for x in Xs]
        for p in range(T):
            rowptr, col = o.csr[p]
            Y = np.zeros((int(o.offsets[-1]), Xs[p].shape[1]))
            rows = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr.astype(np.int64)))
            np.add.at(Y, rows, Xs[p][col.astype(np.int64)])
            for t in range(T):
                V[t] += Y[o.offsets[t]:o.offsets[t + 1]]
        return V

    def epoch(self, n_iters=6):
        o = self.o
        T = o.T
        lr = o.cfg["learning_rate"]
        H = [x.copy() for x in self.X]
        h_t, z, probs = [[None, None] for _ in range(T)], [[None, None] for _ in range(T)], [None] * T
        # forward
        for layer in (0, 1):
            if layer >= n_iters:
                break
            Xs = []
            for p in range(T):
                h_t[p][layer] = H[p]
                x = H[p] @ self.W[p][layer]
                if layer != 0:
                    x = x * self.nrm[p][:, None]
                Xs.append(x)
            V = self.gas(Xs)
            for p in range(T):
                V[p] = V[p] * self.nrm[p][:, None]
                z[p][layer] = V[p]
                if layer == 0:
                    H[p] = np.maximum(V[p], 0)
                else:
                    e = np.exp(V[p] - V[p].max(axis=1, keepdims=True))
                    probs[p] = e / e.sum(axis=1, keepdims=True)
        self.probs = probs
        if n_iters <= 2:
            return
        train = [int(o.n[p] * o.cfg["train_ratio"]) for p in range(T)]
        pmy = []
        for p in range(T):
            y = np.zeros_like(probs[p])
            y[np.arange(o.n[p]), self.labels[p]] = 1.0
            d = probs[p] - y
            d[train[p]:] = 0
            pmy.append(d)
        g = [pmy[p] @ self.W[p][1].T for p in range(T)]                       # iter 2
        V = self.gas([pmy[p] * self.nrm[p][:, None] for p in range(T)])       # iter 3
        for p in range(T):
            v = V[p] * self.nrm[p][:, None]
            d = h_t[p][1].T @ v / train[p]
            self.W[p][1] = self.W[p][1] - lr * d
        avg = sum(self.W[p][1] for p in range(T)) / T
        for p in range(T):
            self.W[p][1] = avg.copy()
        gz = [g[p] * (z[p][0] > 0) for p in range(T)]                         # iter 4
        V = self.gas([gz[p] * self.nrm[p][:, None] for p in range(T)])       # iter 5 (no in-degree scaling)
        for p in range(T):
            d = h_t[p][0].T @ V[p] / train[p]
            self.W[p][0] = self.W[p][0] - lr * d
        avg = sum(self.W[p][0] for p in range(T)) / T
        for p in range(T):
            self.W[p][0] = avg.copy()
