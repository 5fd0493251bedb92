// engine.cpp -- device-resident secure-GCN engine (CoGNN-Opt operators) above the C ABI.  See engine.h.
//
// Reference dataflow being served (file:line in /root/reference):
//   iteration driver            include/ss_vertex_centric_algo_kernel.h:680-910 (ALICE), 912-1189 (BOB)
//   PreScatterComp              algo_kernels/vertex_centric/optimize-gcn/gcn.h:198-255
//   ScatterComp/UpdatePreMerge  gcn.h:257-342  (fused with the OM expand / extract of ssk.h:751-821 into one gather)
//   GatherComp                  gcn.h:375-494
//   ApplyComp                   gcn.h:515-811  (forward, prediction, two-step backward, gradient, FedAvg 747-802)
//   onAlgoKernelStart           gcn.h:819-887  (feature normalisation, Glorot weights with srand(42), weight sharing)
// Protocol (frozen in DESIGN.md): Beaver triples for every multiplication, masked fused gather for Scatter/Gather,
// SecureML local truncation, dealer randomness from the device ChaCha20 PRG.  ReLU / softmax / ReLU' are
// 2PC-RESIDUAL: they stay on the reference's MPC backend; here a clearly labelled ideal-functionality stand-in on
// the HOST lets an epoch run end to end (not secure, not accelerated).
#include "engine.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include <stdexcept>
#include <thread>

namespace cognn {

// true while the online phase of an iteration is being recorded into a CUDA graph: nothing may allocate or synchronise
static thread_local bool g_capturing = false;


// ------------------------------------------------------------------------------------------------------------------
// errors: the reference prints and exit(-1)s (ssk.h:794-797); library code throws, the C shim maps to exit(-1)
// ------------------------------------------------------------------------------------------------------------------
static void ck(cgb_ctx* ctx, int rc, const char* what) {
    if (rc != CGB_OK) throw std::runtime_error(std::string(what) + ": " + cgb_last_error(ctx));
}

bool GNNConfig::read(const std::string& file, std::string* err) {
    std::ifstream fin(file);
    if (!fin.is_open()) {
        if (err) *err = "Failed to open the file: " + file;
        return false;
    }
    std::string param;
    char colon;
    while (fin >> param >> colon) {  // task.h:123-160: "<name> : <value>"
        if (colon != ':') {
            if (err) *err = "Invalid format: expected a colon after " + param;
            return false;
        }
        if (param == "num_layers") fin >> num_layers;
        else if (param == "num_labels") fin >> num_labels;
        else if (param == "input_dim") fin >> input_dim;
        else if (param == "hidden_dim") fin >> hidden_dim;
        else if (param == "num_samples") fin >> num_samples;
        else if (param == "num_edges") fin >> num_edges;
        else if (param == "learning_rate") fin >> learning_rate;
        else if (param == "train_ratio") fin >> train_ratio;
        else if (param == "val_ratio") fin >> val_ratio;
        else if (param == "test_ratio") fin >> test_ratio;
        else {
            if (err) *err = "Unknown parameter: " + param;
            return false;
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------------------------
// graph tile + index vectors (graph_io_util.h:40-208, graph.h:607-641, ssk.h:295-534 with -r 1)
// ------------------------------------------------------------------------------------------------------------------
PartyGraph build_party_graph(const int64_t* edges, size_t n_edges, const int64_t* tid, size_t n_vertices, int T, int me) {
    PartyGraph g;
    g.T = T;
    g.me = me;
    std::vector<uint32_t> local_index(n_vertices);
    std::vector<uint32_t> count(T, 0);
    for (size_t v = 0; v < n_vertices; ++v) {
        if (tid[v] < 0 || tid[v] >= T) throw std::runtime_error("build_party_graph: tile id out of range");
        local_index[v] = count[tid[v]]++;  // ascending vid inside a party (ssk.h:462-464)
    }
    g.offsets.assign(T + 1, 0);
    for (int t = 0; t < T; ++t) g.offsets[t + 1] = g.offsets[t] + count[t];
    const uint32_t n_local = count[me];
    g.vids.reserve(n_local);
    for (size_t v = 0; v < n_vertices; ++v)
        if (tid[v] == me) g.vids.push_back(v);
    std::vector<uint64_t> in_deg(n_local, 0), local_in(n_local, 0);
    g.is_border.assign(n_local, 0);
    struct E { uint32_t row, col; };
    std::vector<E> mine;
    for (size_t e = 0; e < n_edges; ++e) {
        const int64_t s = edges[2 * e], d = edges[2 * e + 1];
        if (s < 0 || d < 0 || (size_t)s >= n_vertices || (size_t)d >= n_vertices)
            throw std::runtime_error("build_party_graph: vertex id out of range");
        if (tid[d] == me) in_deg[local_index[d]]++;  // graph.h:627-632 (local) and graph_io_util.h:170-175 (remote)
        if (tid[s] == me) {
            mine.push_back({g.offsets[tid[d]] + local_index[d], local_index[s]});
            if (tid[d] == me) local_in[local_index[d]]++;
            else g.is_border[local_index[s]] = 1;
        }
    }
    // rows grouped by destination ascending, sources ascending inside a row (edges are (src,dst)-sorted, graph.h:636-641)
    std::sort(mine.begin(), mine.end(), [](const E& a, const E& b) { return a.row != b.row ? a.row < b.row : a.col < b.col; });
    const uint32_t n_rows = g.offsets[T];
    g.rowptr.assign((size_t)n_rows + 1, 0);
    g.col.resize(mine.size());
    for (size_t i = 0; i < mine.size(); ++i) {
        g.rowptr[mine[i].row + 1]++;
        g.col[i] = mine[i].col;
    }
    for (uint32_t r = 0; r < n_rows; ++r) g.rowptr[r + 1] += g.rowptr[r];
    g.in_deg_raw = in_deg;
    g.in_deg = in_deg;
    // ssk.h:412-418: a local vertex without LOCAL in-edge gets a dummy self edge and its degrees are incremented; the
    // edge itself carries no value (dropped at Gather via isGatherDstVertexDummy), so only the increment survives here
    for (uint32_t i = 0; i < n_local; ++i)
        if (local_in[i] == 0) g.in_deg[i] += 1;
    return g;
}

PartyGraph build_party_graph_device(cgb_ctx* ctx, const int64_t* edges, size_t n_edges, const int64_t* tid, size_t n_vertices,
                                    int T, int me) {
    cgb_party_graph* pg = nullptr;
    ck(ctx, cgb_party_graph_build_host(ctx, edges, n_edges, tid, n_vertices, T, me, &pg), "cgb_party_graph_build_host");
    PartyGraph g;
    g.T = T;
    g.me = me;
    const uint32_t n = cgb_party_graph_num_local(pg);
    g.offsets.assign(cgb_party_graph_offsets(pg), cgb_party_graph_offsets(pg) + T + 1);
    g.vids.resize(n);
    g.in_deg_raw.resize(n);
    g.in_deg.resize(n);
    g.is_border.resize(n);
    int rc = CGB_OK;
    if (n) {
        rc |= cgb_d2h(ctx, g.vids.data(), cgb_party_graph_vids(pg), (size_t)n * 8);
        rc |= cgb_d2h(ctx, g.in_deg_raw.data(), cgb_party_graph_in_deg_raw(pg), (size_t)n * 8);
        rc |= cgb_d2h(ctx, g.in_deg.data(), cgb_party_graph_in_deg(pg), (size_t)n * 8);
        rc |= cgb_d2h(ctx, g.is_border.data(), cgb_party_graph_is_border(pg), (size_t)n);
    }
    rc |= cgb_ctx_sync(ctx);
    if (rc == CGB_OK) rc = cgb_party_graph_csr(ctx, pg, &g.csr);
    std::string err = rc == CGB_OK ? "" : cgb_last_error(ctx);
    cgb_party_graph_destroy(ctx, pg);
    if (rc != CGB_OK) throw std::runtime_error("build_party_graph_device: " + err);
    return g;
}

// ------------------------------------------------------------------------------------------------------------------
// device matrices
// ------------------------------------------------------------------------------------------------------------------
struct DMat {
    cgb_ctx* ctx = nullptr;
    uint64_t* p = nullptr;
    uint32_t rows = 0, cols = 0;
    size_t cap = 0;
    DMat() {}
    DMat(const DMat&) = delete;
    DMat& operator=(const DMat&) = delete;
    DMat(DMat&& o) noexcept { *this = std::move(o); }
    DMat& operator=(DMat&& o) noexcept {
        if (this != &o) {
            release();
            ctx = o.ctx; p = o.p; rows = o.rows; cols = o.cols; cap = o.cap;
            o.p = nullptr; o.cap = 0; o.rows = o.cols = 0;
        }
        return *this;
    }
    ~DMat() { release(); }
    void release() {
        if (p) cgb_free(ctx, p);
        p = nullptr;
        cap = 0;
    }
    size_t n() const { return (size_t)rows * cols; }
    void resize(cgb_ctx* c, uint32_t r, uint32_t cl) {
        ctx = c;
        const size_t need = (size_t)r * cl;
        if (need > cap) {
            if (g_capturing) throw std::runtime_error("engine: a device buffer grew while an iteration was being captured");
            // buffers may still be read by queued kernels: the stream is in order, and cudaFree synchronises
            release();
            void* q = nullptr;
            ck(c, cgb_malloc(c, std::max<size_t>(need, 2) * sizeof(uint64_t), &q), "cgb_malloc");
            p = (uint64_t*)q;
            cap = need;
        }
        rows = r;
        cols = cl;
    }
    // on the stream of `c` -- the context the CALLER is issuing on (a side's inside for_sides), not the one `o` was last resized by
    void copy_from(cgb_ctx* c, const DMat& o) {
        resize(c, o.rows, o.cols);
        if (n()) ck(ctx, cgb_d2d(ctx, p, o.p, n() * sizeof(uint64_t)), "cgb_d2d");
    }
};

// PRG stream ids (DESIGN.md "Randomness"); must match oracle/epoch.py
enum Kind : uint64_t { K_FEAT = 1, K_WEIGHT, K_OM_R, K_OM_S, K_MM_U0, K_MM_U1, K_MM_V0, K_MM_V1, K_MM_Z0,
                       K_RM_A0, K_RM_A1, K_RM_B0, K_RM_B1, K_RM_C0, K_RESHARE };
// The engine passes the id of iteration (it % 6); the epoch part ((it - it % 6) << 16) sits in a device word that every
// PRG kernel adds (cgb_ctx_set_prg_stream_bias), so that the captured graph of an iteration can be replayed in later
// epochs.  Sum = (kind << 48) | (it << 16) | (owner << 8) | sub as long as it < 2^32.
static uint64_t stream_id(uint64_t kind, uint64_t it, uint64_t owner, uint64_t sub) {
    return (kind << 48) | ((it % 6) << 16) | (owner << 8) | sub;
}

// one share-side of one owner's state: share 0 lives on the owner, share 1 on its primary helper (owner + 1) % T
struct Side {
    int owner = 0, share = 0, peer = 0;
    uint32_t n = 0;
    // a stream (context) of its own: between two exchanges the sides a process hosts are independent, so their kernels form
    // parallel branches of the iteration's CUDA graph instead of one dependent chain (COGNN_B200_BRANCHES=0: one chain)
    cgb_ctx* sctx = nullptr;
    cudaEvent_t done = nullptr;
    DMat X, X_backup, W[2], h_t[2], z[2], g;
    // per-iteration temporaries
    DMat Xp, V, Y, m, tmp, tmp2;
    // dealer correlations of the current iteration (offline phase), indexed by `sub`
    DMat mmU[2], mmV[2], mmZ[2], rmA[2], rmB[2], rmC[2];
    DMat mm_mine, mm_peer, rm_mine, rm_peer;
    DMat res_in, res_plain, res_plain2, P, Ppeer, grad;
    DMat prob;                    // opened predictions as doubles (bit pattern in u64 words), owner side
    double* h_prob = nullptr;     // pinned host copy, read after the iteration's synchronise
    size_t h_prob_cap = 0;
    std::vector<DMat> upd_recv;  // helper side: masked sums received from the other owners
    DMat delta, S, Smask;  // Smask: sum of the OM mask shares this side holds for its own block (dealt offline)
    // dealer emulation temporaries (offline phase): grow-only members, so that phase neither allocates nor synchronises
    // after the first epoch and can be replayed as a CUDA graph like the online phase
    DMat dl_U0, dl_V0, dl_a0, dl_b0, dl_c0, dl_r;
};

struct PartyData {
    PartyGraph g;
    cgb_csr* csr = nullptr;
    DMat norm;  // enc((inDeg+1)^-1/2), private to the owner
    std::vector<int32_t> labels;
    int32_t* d_labels = nullptr;
    std::vector<double> feats;  // normalised, n x F
};

struct SSGcnEngine::Impl {
    Comm* comm;
    cgb_ctx* ctx;             // the context kernels are issued on RIGHT NOW: the main one, or a side's inside for_sides
    cgb_ctx* main_ctx = nullptr;
    bool branches = !(getenv("COGNN_B200_BRANCHES") && atoi(getenv("COGNN_B200_BRANCHES")) == 0);
    cudaEvent_t fork_ev = nullptr;
    cudaEvent_t ev_on[2] = {nullptr, nullptr};  // around the online phase of an iteration (seconds_online_gpu)
    std::vector<cgb_ctx*> side_ctxs;
    uint64_t total_launches() const {
        uint64_t n = cgb_ctx_launch_count(main_ctx);
        for (cgb_ctx* c : side_ctxs) n += cgb_ctx_launch_count(c);
        return n;
    }
    GNNConfig cfg;
    int f;
    uint32_t key[8];
    int T;
    uint32_t F, H, C;
    std::map<int, PartyData> party;
    std::map<int, Side> own, hlp;
    std::vector<uint32_t> n_of;  // vertices per party (global knowledge: the partition file is public)
    uint64_t lr_fixed = 0;

    // epoch part of the PRG stream ids, on the device (see stream_id): written before every iteration
    uint64_t* d_bias = nullptr;
    uint64_t* h_bias = nullptr;  // pinned
    // CUDA graphs of the online phase, one per iteration index in the epoch (it % 6).  Epoch 0 runs eagerly (buffers
    // grow to their final sizes), the first later occurrence of an iteration index is captured, then replayed: the
    // named graphs are launch-bound (300-600 launches of a few microseconds each per epoch).  COGNN_B200_GRAPHS=0
    // switches it off; recording the message transcript or profiling phases needs the eager path as well.
    bool graphs_enabled = !(getenv("COGNN_B200_GRAPHS") && atoi(getenv("COGNN_B200_GRAPHS")) == 0);
    struct IterGraph {
        cudaGraphExec_t exec = nullptr;
        uint64_t launches = 0, words = 0, rounds = 0;
        std::vector<std::pair<uint32_t, uint32_t>> shapes;  // rows x cols of every side tensor after this iteration
    } graph[6], deal_graph[6];
    // the host-side shape of every tensor a side holds: a replayed graph changes the contents but runs no host code, so
    // the shapes recorded when the iteration was captured are put back (device pointers never change after epoch 0)
    template <typename Fn>
    void for_side_mats(Fn&& fn) {
        for_sides([&](Side& s) {
            DMat* mats[] = {&s.X, &s.X_backup, &s.W[0], &s.W[1], &s.h_t[0], &s.h_t[1], &s.z[0], &s.z[1], &s.g, &s.Xp, &s.V,
                            &s.Y, &s.m, &s.tmp, &s.tmp2, &s.P, &s.Ppeer, &s.grad, &s.prob, &s.res_in, &s.res_plain, &s.res_plain2,
                            &s.mmU[0], &s.mmU[1], &s.mmV[0], &s.mmV[1], &s.mmZ[0], &s.mmZ[1], &s.rmA[0], &s.rmA[1], &s.rmB[0],
                            &s.rmB[1], &s.rmC[0], &s.rmC[1], &s.delta, &s.S, &s.Smask, &s.mm_mine, &s.mm_peer, &s.rm_mine, &s.rm_peer};
            for (DMat* m : mats) fn(*m);
        });
    }
    uint64_t replayed_launches = 0, graph_replays = 0, graph_captures = 0;
    std::vector<int> metrics_pending;  // owners whose opened predictions sit in h_prob until the iteration's synchronise

    // per-phase wall time under the reference's print_duration tags (ssk.h:745,765,802,808,822,856,881,897); only when
    // CGB_ENGINE_PROFILE is set, because it synchronises the stream at every tick
    bool profile = getenv("CGB_ENGINE_PROFILE") != nullptr;
    std::map<std::string, double> phase_s;
    std::chrono::high_resolution_clock::time_point last_tick;
    void tick(const char* tag) {
        if (!profile) return;
        cgb_ctx_sync(ctx);
        auto now = std::chrono::high_resolution_clock::now();
        if (tag) phase_s[tag] += std::chrono::duration<double>(now - last_tick).count();
        last_tick = now;
    }

    int q(int p) const { return (p + 1) % T; }
    Side* side(int owner, int share) {
        auto& m = share == 0 ? own : hlp;
        auto it = m.find(owner);
        return it == m.end() ? nullptr : &it->second;
    }

    // ---- thin wrappers ------------------------------------------------------------------------------------------
    void prg(uint64_t kind, uint64_t it, int owner, int sub, DMat& out, uint32_t r, uint32_t c) {
        out.resize(ctx, r, c);
        ck(ctx, cgb_prg_fill(ctx, key, stream_id(kind, it, owner, sub), 0, out.p, out.n()), "cgb_prg_fill");
    }
    void vadd(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { ck(ctx, cgb_add(ctx, a, b, o, n), "cgb_add"); }
    void vsub(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { ck(ctx, cgb_sub(ctx, a, b, o, n), "cgb_sub"); }
    void send(int src, int dst, const DMat& m, const std::string& tag) { comm->post_send(src, dst, m.p, m.n(), tag); }
    void recv(int dst, int src, DMat& m) { comm->post_recv(dst, src, m.p, m.n()); }
    static std::string tagf(const char* base, int sub, int owner) {
        char b[64];
        if (sub >= 0) snprintf(b, sizeof b, "%s%d.o%d", base, sub, owner);
        else snprintf(b, sizeof b, "%s.o%d", base, owner);
        return b;
    }

    // ---- Beaver matmul (sci::twoPartyGCNMatMul, gcn.h:233,665,671,710): prepare -> exchange -> finish --------------
    // offline (dealer emulation): this side's share of the triple (U, V, Z = U V)
    void mm_deal(Side& s, uint64_t it, int sub, uint32_t M, uint32_t K, uint32_t N) {
        auto sid = [&](uint64_t kind) { return stream_id(kind, it, s.owner, sub); };
        s.mmU[sub].resize(ctx, M, K);
        s.mmV[sub].resize(ctx, K, N);
        s.mmZ[sub].resize(ctx, M, N);
        if (s.share == 0) {  // one launch: the three keystreams of this side's triple share
            cgb_prg_seg sg[3] = {{s.mmU[sub].p, nullptr, s.mmU[sub].n(), sid(K_MM_U0), 0, 0},
                                 {s.mmV[sub].p, nullptr, s.mmV[sub].n(), sid(K_MM_V0), 0, 0},
                                 {s.mmZ[sub].p, nullptr, s.mmZ[sub].n(), sid(K_MM_Z0), 0, 0}};
            ck(ctx, cgb_prg_fill_multi(ctx, key, sg, 3), "cgb_prg_fill_multi");
        } else {
            // Z1 = (U0+U1)(V0+V1) - Z0: one launch makes U1, V1 and the opened sums U0+U1, V0+V1; the product; Z0 comes off
            // as a keystream subtraction in place
            DMat &U = s.dl_U0, &V = s.dl_V0;
            U.resize(ctx, M, K);
            V.resize(ctx, K, N);
            cgb_prg_seg sg[2] = {{U.p, s.mmU[sub].p, U.n(), sid(K_MM_U0), sid(K_MM_U1), 1},
                                 {V.p, s.mmV[sub].p, V.n(), sid(K_MM_V0), sid(K_MM_V1), 1}};
            ck(ctx, cgb_prg_fill_multi(ctx, key, sg, 2), "cgb_prg_fill_multi");
            ck(ctx, cgb_matmul(ctx, U.p, V.p, s.mmZ[sub].p, M, K, N, 0, 0), "cgb_matmul(dealer)");
            ck(ctx, cgb_prg_mask_sub(ctx, key, sid(K_MM_Z0), 0, s.mmZ[sub].p, s.mmZ[sub].p, s.mmZ[sub].n()), "cgb_prg_mask_sub");
        }
    }
    // online: [E_i | F_i] = [A - U | B - V] in one message
    void mm_prepare(Side& s, uint64_t, int sub, const DMat& A, const DMat& B) {
        const uint32_t M = A.rows, K = A.cols, N = B.cols;
        if (s.mmU[sub].rows != M || s.mmU[sub].cols != K || s.mmV[sub].cols != N)
            throw std::runtime_error("mm_prepare: triple was dealt for another shape");
        s.mm_mine.resize(ctx, 1, M * K + K * N);
        s.mm_peer.resize(ctx, 1, M * K + K * N);
        ck(ctx, cgb_sub_pair(ctx, A.p, s.mmU[sub].p, (size_t)M * K, B.p, s.mmV[sub].p, (size_t)K * N, s.mm_mine.p), "cgb_sub_pair");
    }
    void mm_post(Side& s, int sub) {
        const int holder = s.share == 0 ? s.owner : q(s.owner);
        const int other = s.share == 0 ? q(s.owner) : s.owner;
        send(holder, other, s.mm_mine, tagf("mm", sub, s.owner));
        recv(holder, other, s.mm_peer);
    }
    void mm_finish(Side& s, int sub, uint32_t M, uint32_t K, uint32_t N, DMat& C_out) {
        C_out.resize(ctx, M, N);  // E | F are opened (mine += peer) by the launch that also forms V + F for share 0
        ck(ctx, cgb_beaver_matmul_finish_open(ctx, s.mm_mine.p, s.mm_peer.p, s.mmU[sub].p, s.mmV[sub].p, s.mmZ[sub].p, C_out.p, M, K,
                                              N, s.share, f), "cgb_beaver_matmul_finish_open");
    }

    // ---- Beaver row scaling (sci::twoPartyGCNVectorScale, gcn.h:247,476): scaler private to the owner -------------
    void rm_deal(Side& s, uint64_t it, int sub, uint32_t rows, uint32_t D) {
        auto sid = [&](uint64_t kind) { return stream_id(kind, it, s.owner, sub); };
        s.rmA[sub].resize(ctx, rows, D);
        s.rmB[sub].resize(ctx, 1, rows);
        s.rmC[sub].resize(ctx, rows, D);
        if (s.share == 0) {
            cgb_prg_seg sg[3] = {{s.rmA[sub].p, nullptr, s.rmA[sub].n(), sid(K_RM_A0), 0, 0},
                                 {s.rmB[sub].p, nullptr, s.rmB[sub].n(), sid(K_RM_B0), 0, 0},
                                 {s.rmC[sub].p, nullptr, s.rmC[sub].n(), sid(K_RM_C0), 0, 0}};
            ck(ctx, cgb_prg_fill_multi(ctx, key, sg, 3), "cgb_prg_fill_multi");
        } else {
            // c1 = (a0+a1) * (b0+b1)[row] - c0: one launch for a1, b1, the opened sums and c0, one for the product
            DMat &a = s.dl_a0, &b = s.dl_b0, &c0 = s.dl_c0;
            a.resize(ctx, rows, D);
            b.resize(ctx, 1, rows);
            c0.resize(ctx, rows, D);
            cgb_prg_seg sg[3] = {{a.p, s.rmA[sub].p, a.n(), sid(K_RM_A0), sid(K_RM_A1), 1},
                                 {b.p, s.rmB[sub].p, b.n(), sid(K_RM_B0), sid(K_RM_B1), 1},
                                 {c0.p, nullptr, c0.n(), sid(K_RM_C0), 0, 0}};
            ck(ctx, cgb_prg_fill_multi(ctx, key, sg, 3), "cgb_prg_fill_multi");
            ck(ctx, cgb_rowmul_sub(ctx, a.p, b.p, c0.p, s.rmC[sub].p, rows, D), "cgb_rowmul_sub");
        }
    }
    void rm_prepare(Side& s, uint64_t, int sub, const DMat& x, const DMat* scaler) {
        const uint32_t rows = x.rows, D = x.cols;
        if (s.rmA[sub].rows != rows || s.rmA[sub].cols != D) throw std::runtime_error("rm_prepare: triple was dealt for another shape");
        s.rm_mine.resize(ctx, 1, rows * D + rows);
        s.rm_peer.resize(ctx, 1, rows * D + rows);
        // [x - a | scaler - b]; the helper's share of the owner-private scaler is zero
        ck(ctx, cgb_sub_pair(ctx, x.p, s.rmA[sub].p, (size_t)rows * D, s.share == 0 ? scaler->p : nullptr, s.rmB[sub].p, rows,
                             s.rm_mine.p), "cgb_sub_pair");
    }
    void rm_post(Side& s, int sub) {
        const int holder = s.share == 0 ? s.owner : q(s.owner);
        const int other = s.share == 0 ? q(s.owner) : s.owner;
        send(holder, other, s.rm_mine, tagf("rm", sub, s.owner));
        recv(holder, other, s.rm_peer);
    }
    void rm_finish(Side& s, int sub, uint32_t rows, uint32_t D, DMat& out) {
        out.resize(ctx, rows, D);
        ck(ctx, cgb_rowmul_beaver_finish_open(ctx, s.rm_mine.p, s.rm_peer.p, s.rmA[sub].p, s.rmB[sub].p, s.rmC[sub].p, out.p, rows,
                                              D, s.share, f), "cgb_rowmul_beaver_finish_open");
    }

    // run `fn(side)` for every locally hosted side in the canonical order (owner 0 share 0, owner 0 share 1, ...)
    template <typename Fn>
    void for_sides(Fn&& fn) {
        if (!branches || !fork_ev || ctx != main_ctx) {  // one chain (also: nested call, or before setup created the streams)
            for (int o = 0; o < T; ++o)
                for (int sh = 0; sh < 2; ++sh)
                    if (Side* s = side(o, sh)) fn(*s);
            return;
        }
        // fork: every side's stream continues after what the main stream has queued so far; join: the main stream (exchanges,
        // weight averaging, the next fork) continues after every side.  Under stream capture this records parallel branches.
        cudaStream_t ms = (cudaStream_t)cgb_ctx_stream(main_ctx);
        if (cudaEventRecord(fork_ev, ms) != cudaSuccess) throw std::runtime_error("for_sides: cudaEventRecord failed");
        for (int o = 0; o < T; ++o)
            for (int sh = 0; sh < 2; ++sh)
                if (Side* s = side(o, sh)) {
                    if (!s->sctx) {  // no stream for this side (should not happen after setup): run it on the main one
                        fn(*s);
                        continue;
                    }
                    cudaStream_t st = (cudaStream_t)cgb_ctx_stream(s->sctx);
                    if (cudaStreamWaitEvent(st, fork_ev, 0) != cudaSuccess) throw std::runtime_error("for_sides: fork failed");
                    ctx = s->sctx;
                    try {
                        fn(*s);
                    } catch (...) {
                        ctx = main_ctx;
                        throw;
                    }
                    ctx = main_ctx;
                    if (cudaEventRecord(s->done, st) != cudaSuccess || cudaStreamWaitEvent(ms, s->done, 0) != cudaSuccess)
                        throw std::runtime_error("for_sides: join failed");
                }
    }

    // a full two-party row scaling of `x` (both sides), result in `out` chosen by the accessor
    template <typename GetX, typename GetOut>
    void rowscale_all(uint64_t it, int sub, GetX&& getx, GetOut&& getout) {
        for_sides([&](Side& s) {
            DMat& x = getx(s);
            rm_prepare(s, it, sub, x, s.share == 0 ? &party[s.owner].norm : nullptr);
            rm_post(s, sub);
        });
        comm->exchange();
        for_sides([&](Side& s) {
            DMat& x = getx(s);
            // rm_finish reads only the opened message and the triple, never x, so the result may land in x itself; buffers
            // are never swapped, which keeps every device pointer stable across iterations (CUDA-graph replays rely on it)
            rm_finish(s, sub, x.rows, x.cols, getout(s));
        });
    }

    // offline: the owner's OM correlation delta = A r - S (S = the mask shares s_{p->t} of all destination blocks)
    void om_deal(Side& s, uint64_t it, uint32_t D) {
        {   // the mask shares this side holds for its destination block, summed: dealer material, so the online GatherComp
            // additions are one plain sum (share 0 holds s_{p->t} of every other source party, share 1 holds s_{t->t})
            const int t = s.owner;
            uint64_t streams[16];
            uint32_t n_st = 0;
            if (T > 16) throw std::runtime_error("om_deal: more than 16 parties");
            if (s.share == 0) {
                for (int p = 0; p < T; ++p)
                    if (p != t) streams[n_st++] = stream_id(K_OM_S, it, p, t);
            } else {
                streams[n_st++] = stream_id(K_OM_S, it, t, t);
            }
            s.Smask.resize(ctx, s.n, D);
            ck(ctx, cgb_prg_sum(ctx, key, streams, n_st, nullptr, 0, s.Smask.p, s.Smask.n()), "cgb_prg_sum(mask shares)");
        }
        if (s.share != 0) return;
        PartyData& pd = party[s.owner];
        const uint32_t n_rows = pd.g.offsets[T];
        DMat& r = s.dl_r;
        r.resize(ctx, s.n, D);
        s.S.resize(ctx, n_rows, D);
        {   // one launch: the input mask r and the T output-mask blocks s_{p->t}
            std::vector<cgb_prg_seg> sg;
            sg.push_back({r.p, nullptr, r.n(), stream_id(K_OM_R, it, s.owner, 0), 0, 0});
            for (int t = 0; t < T; ++t)
                sg.push_back({s.S.p + (size_t)pd.g.offsets[t] * D, nullptr, (size_t)n_of[t] * D, stream_id(K_OM_S, it, s.owner, t), 0, 0});
            ck(ctx, cgb_prg_fill_multi(ctx, key, sg.data(), (uint32_t)sg.size()), "cgb_prg_fill_multi");
        }
        s.delta.resize(ctx, n_rows, D);
        ck(ctx, cgb_gather_sum(ctx, pd.csr, r.p, nullptr, s.delta.p, D), "gather(dealer)");
        vsub(s.delta.p, s.S.p, s.delta.p, s.delta.n());
    }

    // everything the dealer hands out for iteration `it` (shapes per SURVEY 3.4)
    void deal_iteration(uint64_t it) {
        const int ph = (int)(it % 6);
        for_sides([&](Side& s) {
            const uint32_t n = s.n;
            switch (ph) {
                case 0: mm_deal(s, it, 0, n, F, H); om_deal(s, it, H); rm_deal(s, it, 1, n, H); break;
                case 1: mm_deal(s, it, 0, n, H, C); rm_deal(s, it, 0, n, C); om_deal(s, it, C); rm_deal(s, it, 1, n, C); break;
                case 2: mm_deal(s, it, 0, n, C, H); break;
                case 3: rm_deal(s, it, 0, n, C); om_deal(s, it, C); rm_deal(s, it, 1, n, C); mm_deal(s, it, 1, H, n, C); break;
                case 4: break;
                case 5: rm_deal(s, it, 0, n, H); om_deal(s, it, H); mm_deal(s, it, 1, F, n, H); break;
            }
        });
    }

    // ---- Scatter / Gather of one GAS iteration: V = Xp + sum over in-edges (ssk.h:748-880 share-local composite) ---
    void gas(uint64_t it) {
        // round 1: helper -> owner OM online message m = Xp1 - r
        for_sides([&](Side& s) {
            const uint32_t D = s.Xp.cols;
            if (s.share == 1) {
                s.m.resize(ctx, s.n, D);
                ck(ctx, cgb_prg_mask_sub(ctx, key, stream_id(K_OM_R, it, s.owner, 0), 0, s.Xp.p, s.m.p, s.m.n()), "mask");
                send(q(s.owner), s.owner, s.m, tagf("om", -1, s.owner));
            } else {
                s.m.resize(ctx, s.n, D);
                recv(s.owner, q(s.owner), s.m);
            }
        });
        comm->exchange();
        // owner: fused gather over all its out-edges, one block of rows per destination party
        for_sides([&](Side& s) {
            if (s.share != 0) return;
            PartyData& pd = party[s.owner];
            const uint32_t D = s.Xp.cols, n_rows = pd.g.offsets[T];
            // online: Y = A (Xp0 + m) + delta
            vadd(s.Xp.p, s.m.p, s.m.p, s.m.n());
            s.Y.resize(ctx, n_rows, D);
            if (s.delta.cols != D) throw std::runtime_error("gas: OM correlation was dealt for another width");
            ck(ctx, cgb_gather_sum(ctx, pd.csr, s.m.p, s.delta.p, s.Y.p, D), "cgb_gather_sum");
        });
        // round 2: mirror-update blocks to the primary helper of each destination owner (ssk.h:1090)
        for (int p = 0; p < T; ++p) {
            for (int t = 0; t < T; ++t) {
                if (t == p || q(t) == p) continue;
                if (Side* s = side(p, 0)) {
                    const uint32_t D = s->Xp.cols;
                    char tg[64];
                    snprintf(tg, sizeof tg, "upd.o%d.t%d", p, t);
                    comm->post_send(p, q(t), s->Y.p + (size_t)party[p].g.offsets[t] * D, (size_t)n_of[t] * D, tg);
                }
                if (Side* h = side(t, 1)) {
                    const uint32_t D = h->Xp.cols;
                    if (h->upd_recv.size() < (size_t)T) h->upd_recv.resize(T);
                    h->upd_recv[p].resize(ctx, h->n, D);
                    comm->post_recv(q(t), p, h->upd_recv[p].p, h->upd_recv[p].n());
                }
            }
        }
        comm->exchange();
        // GatherComp additions (gcn.h:456-463)
        for_sides([&](Side& s) {
            const uint32_t D = s.Xp.cols;
            const int t = s.owner;
            s.V.resize(ctx, s.n, D);
            // one launch: V = Xp + (the blocks this side holds) + (the sum of the mask shares it holds, dealt offline)
            const uint64_t* in[16];
            uint32_t n_in = 0;
            if (T > 14) throw std::runtime_error("gas: more than 14 parties");
            if (s.Smask.rows != s.n || s.Smask.cols != D) throw std::runtime_error("gas: OM mask shares were dealt for another shape");
            in[n_in++] = s.Xp.p;
            in[n_in++] = s.Smask.p;
            if (s.share == 0) {
                in[n_in++] = s.Y.p + (size_t)party[t].g.offsets[t] * D;
            } else {
                for (int p = 0; p < T; ++p) {
                    if (p == t) continue;
                    if (q(t) == p) in[n_in++] = own.at(p).Y.p + (size_t)party[p].g.offsets[t] * D;  // computed on this very party
                    else in[n_in++] = s.upd_recv[p].p;
                }
            }
            ck(ctx, cgb_sum_n(ctx, in, n_in, s.V.p, s.V.n()), "cgb_sum_n");
        });
    }

    // ---- 2PC-RESIDUAL stand-in (ideal functionality; NOT secure, NOT the reference's protocol) ----------------------------
    // helper sends its shares; the owner evaluates the function on the reconstructed values ON THE DEVICE (cgb_ideal_*:
    // exact integer ReLU / ReLU', softmax in IEEE doubles with a restated exp) and re-shares: the helper's new share is a
    // PRG stream both know from the dealer, the owner's is fn(x) - PRG.  kind: 1 = ReLU, 2 = ReLU' mask, 3 = softmax, p - y.
    void residual(uint64_t it, int kind, int n_in, int n_out, DMat* (*in_sel)(Side&, int), DMat* (*out_sel)(Side&, int)) {
        for_sides([&](Side& s) {
            size_t total = 0;
            for (int i = 0; i < n_in; ++i) total += in_sel(s, i)->n();
            s.res_in.resize(ctx, 1, (uint32_t)total);
            if (s.share == 1 && n_in == 1) {  // a single operand goes out from where it lives
                send(q(s.owner), s.owner, *in_sel(s, 0), tagf("res", 0, s.owner));
            } else if (s.share == 1) {
                size_t off = 0;
                for (int i = 0; i < n_in; ++i) {
                    DMat* m = in_sel(s, i);
                    ck(ctx, cgb_d2d(ctx, s.res_in.p + off, m->p, m->n() * 8), "d2d");
                    off += m->n();
                }
                send(q(s.owner), s.owner, s.res_in, tagf("res", 0, s.owner));
            } else {
                recv(s.owner, q(s.owner), s.res_in);
            }
        });
        comm->exchange();
        for_sides([&](Side& s) {
            if (s.share == 0) {
                DMat* a = in_sel(s, 0);
                DMat* plain[2] = {&s.res_plain, &s.res_plain2};
                for (int k = 0; k < n_out; ++k) plain[k]->resize(ctx, a->rows, a->cols);
                if (kind == 1 || kind == 2) {  // the stand-in and the re-sharing of its single result in one launch
                    DMat* o = out_sel(s, 0);
                    o->resize(ctx, a->rows, a->cols);
                    const uint64_t* z0 = kind == 2 ? in_sel(s, 1)->p : nullptr;
                    const uint64_t* z1 = kind == 2 ? s.res_in.p + a->n() : nullptr;
                    ck(ctx, cgb_ideal_relu_reshare(ctx, key, stream_id(K_RESHARE, it, s.owner, 0), a->p, s.res_in.p, z0, z1, o->p,
                                                   a->n()), "cgb_ideal_relu_reshare");
                    return;
                } else {
                    const uint64_t train = (uint64_t)(s.n * cfg.train_ratio);  // gcn.h:560
                    ck(ctx, cgb_ideal_softmax(ctx, a->p, s.res_in.p, party.at(s.owner).d_labels, a->rows, a->cols, train, f,
                                              s.res_plain.p, s.res_plain2.p), "cgb_ideal_softmax");
                }
                for (int k = 0; k < n_out; ++k) {
                    DMat* o = out_sel(s, k);
                    o->resize(ctx, a->rows, a->cols);
                    ck(ctx, cgb_prg_mask_sub(ctx, key, stream_id(K_RESHARE, it, s.owner, k), 0, plain[k]->p, o->p, o->n()), "reshare");
                }
            } else {
                for (int k = 0; k < n_out; ++k) {
                    DMat* o = out_sel(s, k);
                    ck(ctx, cgb_prg_fill(ctx, key, stream_id(K_RESHARE, it, s.owner, k), 0, o->p, o->n()), "reshare");
                }
            }
        });
    }

    // ---- weight averaging (gcn.h:747-802): reduce to parties 0 and 1, public scale 1/T, redistribute ---------------
    // receive buffers and the averages live in the sides (grow-only), so the online phase never allocates
    std::map<int, DMat> wa_rx_own[2], wa_rx_hlp[2];
    DMat wa_A0[2], wa_A1[2];
    void weight_average(uint64_t, int layer) {
        if (T == 1) return;
        char tg[64];
        auto& rx_own = wa_rx_own[layer];  // at party 1: W0_i of i >= 2
        auto& rx_hlp = wa_rx_hlp[layer];  // at party 0: W1_{i-1}
        for (int i = 2; i < T; ++i) {
            if (Side* s = side(i, 0)) {
                snprintf(tg, sizeof tg, "w%d.own%d", layer, i);
                comm->post_send(i, 1, s->W[layer].p, s->W[layer].n(), tg);
            }
            if (comm->is_local(1)) {
                rx_own[i].resize(ctx, own.at(1).W[layer].rows, own.at(1).W[layer].cols);
                comm->post_recv(1, i, rx_own[i].p, rx_own[i].n());
            }
            if (Side* h = side(i - 1, 1)) {  // lives on party i
                snprintf(tg, sizeof tg, "w%d.hlp%d", layer, i - 1);
                comm->post_send(i, 0, h->W[layer].p, h->W[layer].n(), tg);
            }
            if (comm->is_local(0)) {
                rx_hlp[i].resize(ctx, own.at(0).W[layer].rows, own.at(0).W[layer].cols);
                comm->post_recv(0, i, rx_hlp[i].p, rx_hlp[i].n());
            }
        }
        comm->exchange();
        const uint64_t c = (uint64_t)(int64_t)((1.0 / T) * (double)(1ull << f));  // gcn.h:763-764
        DMat& A0 = wa_A0[layer];
        DMat& A1 = wa_A1[layer];
        // parties 0 and 1: sum of the replicas' shares, public scale, and the result written straight over the local and the
        // remote weight copy it replaces (gcn.h:765-777) -- one launch each; the separate buffer only when it is sent on
        auto average = [&](int p, DMat& A, std::map<int, DMat>& rx, int share) {
            DMat& w = own.at(p).W[layer];
            DMat& rw = hlp.at((p - 1 + T) % T).W[layer];  // this party's remoteWeight: share 1 of party p-1's replica
            const uint64_t* in[16];
            uint64_t* out[4];
            uint32_t n_in = 0, n_out = 0;
            if (T > 16) throw std::runtime_error("weight_average: more than 16 parties");
            in[n_in++] = w.p;
            for (int i = 2; i < T; ++i) in[n_in++] = rx[i].p;
            in[n_in++] = rw.p;
            if (T > 2) {
                A.resize(ctx, w.rows, w.cols);
                out[n_out++] = A.p;
            }
            out[n_out++] = w.p;
            out[n_out++] = rw.p;
            ck(ctx, cgb_avg_public(ctx, in, n_in, c, out, n_out, w.n(), f, share), "cgb_avg_public");
        };
        if (comm->is_local(0)) average(0, A0, rx_hlp, 0);
        if (comm->is_local(1)) average(1, A1, rx_own, 1);
        // parties i >= 2 receive both averages where they are used: local weight <- A1, remote weight <- A0
        for (int i = 2; i < T; ++i) {
            snprintf(tg, sizeof tg, "wavg%d.A1", layer);
            if (comm->is_local(1)) comm->post_send(1, i, A1.p, A1.n(), tg);
            snprintf(tg, sizeof tg, "wavg%d.A0", layer);
            if (comm->is_local(0)) comm->post_send(0, i, A0.p, A0.n(), tg);
            if (comm->is_local(i)) {
                DMat& w = own.at(i).W[layer];
                DMat& rw = hlp.at(i - 1).W[layer];
                comm->post_recv(i, 1, w.p, w.n());
                comm->post_recv(i, 0, rw.p, rw.n());
            }
        }
        comm->exchange();
    }
};

// ------------------------------------------------------------------------------------------------------------------
SSGcnEngine::SSGcnEngine(Comm* comm, const GNNConfig& cfg, int f, const uint32_t key[8]) {
    impl_ = new Impl();
    impl_->comm = comm;
    impl_->ctx = impl_->main_ctx = comm->ctx();
    impl_->cfg = cfg;
    impl_->f = f;
    memcpy(impl_->key, key, sizeof(impl_->key));
    impl_->T = comm->world();
    impl_->F = cfg.input_dim;
    impl_->H = cfg.hidden_dim;
    impl_->C = cfg.num_labels;
    if (cfg.num_layers != 2) throw std::runtime_error("SSGcnEngine: the reference operators hard-code 2 layers (gcn.h:898-927)");
    if (f <= 0 || f >= 31) throw std::runtime_error("SSGcnEngine: SCALER_BIT_LENGTH must be in (0, 31) (gcn.h:191)");
    cgb_ctx* ctx = impl_->main_ctx;
    void* p = nullptr;
    ck(ctx, cgb_malloc(ctx, 8, &p), "cgb_malloc");
    impl_->d_bias = (uint64_t*)p;
    if (cgb_host_alloc(8, &p) != CGB_OK) throw std::runtime_error("SSGcnEngine: pinned allocation failed");
    impl_->h_bias = (uint64_t*)p;
    *impl_->h_bias = 0;
    ck(ctx, cgb_memset(ctx, impl_->d_bias, 0, 8), "memset");
    ck(ctx, cgb_ctx_set_prg_stream_bias(ctx, impl_->d_bias), "cgb_ctx_set_prg_stream_bias");
    ck(ctx, cgb_ctx_sync(ctx), "sync");
}

SSGcnEngine::~SSGcnEngine() {
    if (!impl_) return;
    cgb_ctx_sync(impl_->ctx);
    if (impl_->profile)
        for (auto& kv : impl_->phase_s) fprintf(stderr, "::%s took %lf seconds (all iterations)\n", kv.first.c_str(), kv.second);
    for (auto& g : impl_->graph)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& g : impl_->deal_graph)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    cgb_ctx_set_prg_stream_bias(impl_->ctx, nullptr);
    cgb_free(impl_->ctx, impl_->d_bias);
    cgb_host_free(impl_->h_bias);
    for (auto& kv : impl_->own)
        if (kv.second.h_prob) cgb_host_free(kv.second.h_prob);
    for (auto& kv : impl_->party) {
        if (kv.second.csr) cgb_csr_destroy(impl_->ctx, kv.second.csr);
        if (kv.second.d_labels) cgb_free(impl_->ctx, kv.second.d_labels);
    }
    // the sides' buffers were allocated through (and remember) their side contexts: release them first
    for (cgb_ctx* c : impl_->side_ctxs) cgb_ctx_sync(c);
    for (auto* m : {&impl_->own, &impl_->hlp})
        for (auto& kv : *m)
            if (kv.second.done) cudaEventDestroy(kv.second.done);
    impl_->own.clear();
    impl_->hlp.clear();
    for (int l = 0; l < 2; ++l) {
        impl_->wa_rx_own[l].clear(); impl_->wa_rx_hlp[l].clear();
    }
    for (cgb_ctx* c : impl_->side_ctxs) cgb_ctx_destroy(c);
    if (impl_->fork_ev) cudaEventDestroy(impl_->fork_ev);
    for (cudaEvent_t e : impl_->ev_on)
        if (e) cudaEventDestroy(e);
    delete impl_;
}

void SSGcnEngine::add_party(const PartyGraph& g, const double* feats_local, const int32_t* labels_local) {
    Impl& im = *impl_;
    if (!im.comm->is_local(g.me)) throw std::runtime_error("add_party: party is not hosted by this process");
    PartyData& pd = im.party[g.me];
    pd.g = g;
    const uint32_t n = (uint32_t)g.vids.size();
    if (im.n_of.empty()) {
        im.n_of.resize(im.T);
        for (int t = 0; t < im.T; ++t) im.n_of[t] = g.offsets[t + 1] - g.offsets[t];
    }
    pd.labels.assign(labels_local, labels_local + n);
    // gcn.h:819-835, 857-862: raw * (inDeg + 1)^-1/2 with the in-degree BEFORE the dummy increment (ssk.h:177 < 190)
    pd.feats.resize((size_t)n * im.F);
    for (uint32_t i = 0; i < n; ++i) {
        const double sc = std::pow((double)g.in_deg_raw[i] + 1.0, -0.5);
        for (uint32_t j = 0; j < im.F; ++j) pd.feats[(size_t)i * im.F + j] = feats_local[(size_t)i * im.F + j] * sc;
    }
    if (g.csr) {  // device ingest: the CSR is resident already
        pd.csr = g.csr;
        pd.g.csr = nullptr;
    } else {
        ck(im.ctx, cgb_csr_create(im.ctx, g.rowptr.data(), g.col.data(), g.offsets[im.T], g.col.size(), n, &pd.csr), "cgb_csr_create");
    }
    // normaliser (gcn.h:219-221): deg == 0 ? 0 : enc((deg+1)^-1/2); PreScatter is handed inDeg as well (ssk.h:739)
    std::vector<uint64_t> nv(n);
    for (uint32_t i = 0; i < n; ++i)
        nv[i] = g.in_deg[i] == 0 ? 0 : (uint64_t)(int64_t)(std::pow((double)g.in_deg[i] + 1.0, -0.5) * (double)(1ull << im.f));
    pd.norm.resize(im.ctx, 1, n);
    ck(im.ctx, cgb_h2d(im.ctx, pd.norm.p, nv.data(), n * 8), "h2d");
    void* dl = nullptr;  // labels on the device: the prediction-layer stand-in (cgb_ideal_softmax) forms p - y there
    ck(im.ctx, cgb_malloc(im.ctx, std::max<size_t>(n, 1) * sizeof(int32_t), &dl), "cgb_malloc");
    pd.d_labels = (int32_t*)dl;
    if (n) ck(im.ctx, cgb_h2d(im.ctx, pd.d_labels, pd.labels.data(), n * sizeof(int32_t)), "h2d");
    ck(im.ctx, cgb_ctx_sync(im.ctx), "sync");
}

// gcn.h:838-852: std::srand(42) per call, then (double) rand() / RAND_MAX * 2 * limit - limit
static std::vector<double> init_weight(int d0, int d1) {
    std::vector<double> W((size_t)d0 * d1);
    std::srand(42);
    const double limit = std::sqrt(6.0 / (d0 + d1));
    for (int i = 0; i < d0; ++i)
        for (int j = 0; j < d1; ++j) W[(size_t)i * d1 + j] = (double)std::rand() / RAND_MAX * 2 * limit - limit;
    return W;
}

void SSGcnEngine::setup() {
    Impl& im = *impl_;
    cgb_ctx*& ctx = im.ctx;  // follows the side contexts inside for_sides
    const int T = im.T;
    const uint32_t dims[3] = {im.F, im.H, im.C};
    std::vector<double> Wp[2] = {init_weight(im.F, im.H), init_weight(im.H, im.C)};
    im.lr_fixed = (uint64_t)(int64_t)(im.cfg.learning_rate * (double)(1ull << im.f));  // gcn.h:678
    // Every party splits its own features / weight replica (ssk.h:205, gcn.h:880-882) and ships share 1 to its helper
    // (ssk.h:209-232).  With dealer randomness both sides of the split are PRG-reproducible, so the owner computes
    // share 0 and the helper receives share 1 as a message.
    for (int o = 0; o < T; ++o) {
        const int qo = im.q(o);
        if (im.comm->is_local(o)) {
            Side& s = im.own[o];
            s.owner = o; s.share = 0; s.n = im.n_of[o];
            PartyData& pd = im.party.at(o);
            double* dptr = nullptr;
            void* dv = nullptr;
            ck(ctx, cgb_malloc(ctx, std::max<size_t>(pd.feats.size(), 1) * 8, &dv), "malloc");
            dptr = (double*)dv;
            ck(ctx, cgb_h2d(ctx, dptr, pd.feats.data(), pd.feats.size() * 8), "h2d");
            s.X.resize(ctx, s.n, im.F);
            s.tmp.resize(ctx, s.n, im.F);
            ck(ctx, cgb_share_split(ctx, dptr, s.X.n(), im.f, im.key, stream_id(K_FEAT, 0, o, 0), 0, s.X.p, s.tmp.p), "share_split");
            im.comm->post_send(o, qo, s.tmp.p, s.tmp.n(), Impl::tagf("setupX", -1, o));
            ck(ctx, cgb_ctx_sync(ctx), "sync");
            cgb_free(ctx, dv);
        }
        if (im.comm->is_local(qo)) {
            Side& h = im.hlp[o];
            h.owner = o; h.share = 1; h.n = im.n_of[o];
            h.X.resize(ctx, h.n, im.F);
            im.comm->post_recv(qo, o, h.X.p, h.X.n());
        }
    }
    im.comm->exchange();
    for (int l = 0; l < 2; ++l) {
        std::map<int, DMat> w1;
        for (int o = 0; o < T; ++o) {
            const int qo = im.q(o);
            if (im.comm->is_local(o)) {
                Side& s = im.own.at(o);
                void* dv = nullptr;
                ck(ctx, cgb_malloc(ctx, Wp[l].size() * 8, &dv), "malloc");
                ck(ctx, cgb_h2d(ctx, dv, Wp[l].data(), Wp[l].size() * 8), "h2d");
                s.W[l].resize(ctx, dims[l], dims[l + 1]);
                w1[o].resize(ctx, dims[l], dims[l + 1]);
                ck(ctx, cgb_share_split(ctx, (const double*)dv, s.W[l].n(), im.f, im.key, stream_id(K_WEIGHT, 0, o, l), 0,
                                        s.W[l].p, w1[o].p), "share_split");
                im.comm->post_send(o, qo, w1[o].p, w1[o].n(), Impl::tagf("setupW", l, o));
                ck(ctx, cgb_ctx_sync(ctx), "sync");
                cgb_free(ctx, dv);
            }
            if (im.comm->is_local(qo)) {
                Side& h = im.hlp.at(o);
                h.W[l].resize(ctx, dims[l], dims[l + 1]);
                im.comm->post_recv(qo, o, h.W[l].p, h.W[l].n());
            }
        }
        im.comm->exchange();  // ssk.h:231-232
    }
    if (im.branches && cgb_ctx_stream(im.main_ctx) != nullptr) {
        int dev = 0;
        cudaGetDevice(&dev);
        im.for_sides([&](Side& s) {  // (still sequential here: fork_ev does not exist yet)
            if (cgb_ctx_create(dev, &s.sctx) != CGB_OK) throw std::runtime_error("setup: side context creation failed");
            ck(s.sctx, cgb_ctx_set_prg_stream_bias(s.sctx, im.d_bias), "cgb_ctx_set_prg_stream_bias");
            if (cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) throw std::runtime_error("setup: event creation failed");
            im.side_ctxs.push_back(s.sctx);
        });
        if (cudaEventCreateWithFlags(&im.fork_ev, cudaEventDisableTiming) != cudaSuccess) throw std::runtime_error("setup: event creation failed");
    }
    im.for_sides([&](Side& s) {
        s.X_backup.copy_from(ctx, s.X);  // ssk.h:226-227
        s.h_t[0].resize(ctx, im.F, s.n);  // gcn.h:230-231 for layer 0: the transpose of a constant
        ck(ctx, cgb_transpose(ctx, s.X_backup.p, s.h_t[0].p, s.n, im.F), "transpose");
    });
    ck(im.main_ctx, cgb_ctx_sync(im.main_ctx), "sync");
    {   // the most words one party sends another in one online round (two sides may share a pair: x 2), for the plane's arenas
        size_t nmax = 0;
        for (uint32_t v : im.n_of) nmax = std::max<size_t>(nmax, v);
        const size_t F = im.F, H = im.H, C = im.C, D = std::max(H, C);
        size_t m = nmax * F + std::max(F * H, nmax * H);            // Beaver product messages [E | F]
        m = std::max(m, nmax * D + std::max(D * D, nmax * D));
        m = std::max(m, 2 * nmax * D + nmax);                        // row scaling, residual operands, update blocks
        im.comm->reserve(2 * m + 64);
    }
}

static DMat* sel_z0(Side& s, int) { return &s.z[0]; }
static DMat* sel_z1(Side& s, int) { return &s.z[1]; }
static DMat* sel_X(Side& s, int) { return &s.X; }
static DMat* sel_XP(Side& s, int k) { return k == 0 ? &s.P : &s.X; }
static DMat* sel_Xz0(Side& s, int k) { return k == 0 ? &s.X : &s.z[0]; }

// online phase of iteration `it` (everything after the dealer hand-out): only stream-ordered work, no allocation once the
// buffers have their final sizes, no host synchronisation -- so it can be recorded into a CUDA graph
static void online_iteration(SSGcnEngine::Impl& im, uint64_t it) {
    cgb_ctx*& ctx = im.ctx;  // follows the side contexts inside for_sides
    const uint32_t F = im.F, H = im.H, C = im.C;
    const int ph = (int)(it % 6);
    // ssk.h:695, 938: every epoch restarts from the input features.  They never change, so layer 0 reads X_backup (and its
    // transpose, made once in setup) instead of copying it into X first.
    im.tick(nullptr);

    if (ph == 0 || ph == 1) {
        // ---------------- forward layer `ph` ----------------
        const int layer = ph;
        const uint32_t Din = layer == 0 ? F : H, Dout = layer == 0 ? H : C;
        im.for_sides([&](Side& s) {  // PreScatterComp (gcn.h:198-255)
            if (layer != 0) {
                s.h_t[layer].resize(ctx, Din, s.n);
                ck(ctx, cgb_transpose(ctx, s.X.p, s.h_t[layer].p, s.n, Din), "transpose");  // gcn.h:230-231
            }
            im.mm_prepare(s, it, 0, layer == 0 ? s.X_backup : s.X, s.W[layer]);
            im.mm_post(s, 0);
        });
        im.comm->exchange();
        im.for_sides([&](Side& s) { im.mm_finish(s, 0, s.n, Din, Dout, s.Xp); });
        if (layer != 0)  // gcn.h:243-254
            im.rowscale_all(it, 0, [](Side& s) -> DMat& { return s.Xp; }, [](Side& s) -> DMat& { return s.Xp; });
        im.tick("PreScatterComp");
        im.gas(it);
        im.tick("Scatter+Gather (OM message, fused gather-sum, update exchange, adds)");
        // gcn.h:470-484: in-degree scaling ((it + 1) % 6 != 0 for the forward layers)
        // the scaled aggregate is the layer's pre-activation z (gcn.h:546,559): written where the backward pass reads it
        im.rowscale_all(it, 1, [](Side& s) -> DMat& { return s.V; }, [layer](Side& s) -> DMat& { return s.z[layer]; });
        im.tick("Gather_computation (in-degree scale)");
        if (layer == 0) {  // gcn.h:546-558: ReLU -- 2PC-RESIDUAL
            im.for_sides([&](Side& s) { s.X.resize(ctx, s.n, Dout); });
            im.residual(it, 1, 1, 1, sel_z0, sel_X);
        } else {  // gcn.h:559-642: softmax, p - y -- 2PC-RESIDUAL; then p is opened to the owner (gcn.h:604)
            im.for_sides([&](Side& s) {
                s.X.resize(ctx, s.n, Dout);
                s.P.resize(ctx, s.n, Dout);
            });
            im.residual(it, 3, 1, 2, sel_z1, sel_XP);
            im.for_sides([&](Side& s) {
                if (s.share == 1) im.send(im.q(s.owner), s.owner, s.P, SSGcnEngine::Impl::tagf("open_p", -1, s.owner));
                else {
                    s.Ppeer.resize(ctx, s.n, Dout);
                    im.recv(s.owner, im.q(s.owner), s.Ppeer);
                }
            });
            im.comm->exchange();
            im.for_sides([&](Side& s) {  // getPlainShareVecVec (gcn.h:604): decode on the device, copy to pinned host memory
                if (s.share != 0) return;
                // loss / accuracy on the device (cgb_prediction_metrics): a few hundred block records come back, not n x C doubles
                uint32_t nb = 0;
                ck(ctx, cgb_prediction_metrics(ctx, nullptr, nullptr, nullptr, s.n, Dout, 0, 0, im.f, nullptr, &nb), "metrics size");
                s.prob.resize(ctx, std::max<uint32_t>(nb, 1), 4);
                if (s.h_prob_cap < s.prob.n()) {
                    if (g_capturing) throw std::runtime_error("engine: host buffer grew while an iteration was being captured");
                    if (s.h_prob) cgb_host_free(s.h_prob);
                    void* hp = nullptr;
                    if (cgb_host_alloc(std::max<size_t>(s.prob.n(), 1) * 8, &hp) != CGB_OK) throw std::runtime_error("cgb_host_alloc");
                    s.h_prob = (double*)hp;
                    s.h_prob_cap = s.prob.n();
                }
                const uint32_t train = (uint32_t)(s.n * im.cfg.train_ratio), val = (uint32_t)(s.n * im.cfg.val_ratio);
                ck(ctx, cgb_prediction_metrics(ctx, s.P.p, s.Ppeer.p, im.party.at(s.owner).d_labels, s.n, Dout, train, val, im.f,
                                               (double*)s.prob.p, &nb), "cgb_prediction_metrics");
                if (s.n) ck(ctx, cgb_d2h(ctx, s.h_prob, s.prob.p, (size_t)nb * 4 * 8), "d2h");
                im.metrics_pending.push_back(s.owner);
            });
        }
    } else if (ph == 2) {
        // ---------------- backward, first step of the last layer: apply only (ssk.h:709-732; gcn.h:664-669) -----
        im.for_sides([&](Side& s) {
            s.tmp2.resize(ctx, C, H);
            ck(ctx, cgb_transpose(ctx, s.W[1].p, s.tmp2.p, H, C), "transpose");  // weightT (gcn.h:648)
            im.mm_prepare(s, it, 0, s.X, s.tmp2);
            im.mm_post(s, 0);
        });
        im.comm->exchange();
        im.for_sides([&](Side& s) { im.mm_finish(s, 0, s.n, C, H, s.g); });
    } else if (ph == 3 || ph == 5) {
        // ---------------- backward GAS + weight gradient (gcn.h:247-254, 470-484, 671-684 / 710-736) -------------
        const int layer = ph == 3 ? 1 : 0;
        const uint32_t Din = layer == 0 ? F : H, Dout = layer == 0 ? H : C;
        im.rowscale_all(it, 0, [](Side& s) -> DMat& { return s.X; }, [](Side& s) -> DMat& { return s.Xp; });
        im.tick("PreScatterComp");
        im.gas(it);
        im.tick("Scatter+Gather (OM message, fused gather-sum, update exchange, adds)");
        if ((it + 1) % 6 != 0)  // gcn.h:470: no in-degree scaling on the last iteration of the epoch
            im.rowscale_all(it, 1, [](Side& s) -> DMat& { return s.V; }, [](Side& s) -> DMat& { return s.V; });
        im.tick("Gather_computation (in-degree scale)");
        im.for_sides([&](Side& s) {
            im.mm_prepare(s, it, 1, s.h_t[layer], s.V);  // d = h_t * v
            im.mm_post(s, 1);
        });
        im.comm->exchange();
        im.for_sides([&](Side& s) {
            DMat& d = s.grad;
            im.mm_finish(s, 1, Din, s.n, Dout, d);
            const uint64_t train = (uint64_t)(s.n * im.cfg.train_ratio);
            const uint64_t gs = train ? (uint64_t)(int64_t)((1.0 / (double)train) * (double)(1ull << im.f)) : 0;  // gcn.h:673-676
            ck(ctx, cgb_scale_apply_gradient(ctx, s.W[layer].p, d.p, gs, im.lr_fixed, d.p, s.W[layer].p, d.n(), im.f, s.share),
               "cgb_scale_apply_gradient");
            if (layer == 1) s.X.copy_from(ctx, s.g);  // dstVec.swap(g) (gcn.h:684)
            else s.X.copy_from(ctx, s.V);             // first layer: g is empty in the reference; never used again
        });
        im.tick("Apply_computation");
        im.weight_average(it, layer);
        im.tick("Apply_computation (weight averaging)");
    } else {
        // ---------------- ph == 4: ReLU' mask, first layer so no further matmul (gcn.h:702-708) -- 2PC-RESIDUAL ---
        im.residual(it, 2, 2, 1, sel_Xz0, sel_X);
    }
    im.tick("Apply_computation");
}

// loss / accuracy the owner prints after the prediction layer (gcn.h:603-632), from the opened probabilities
static Metrics owner_metrics(SSGcnEngine::Impl& im, uint64_t it, int owner, bool verbose) {
    Side& s = im.own.at(owner);
    const uint32_t n = s.n;
    const uint64_t train = (uint64_t)(n * im.cfg.train_ratio), val = (uint64_t)(n * im.cfg.val_ratio);
    // block records {loss, hits full, hits train, hits test} of cgb_prediction_metrics, added in block order
    const uint32_t nb = (n + 255) / 256;
    double loss = 0, hit_full = 0, hit_train = 0, hit_test = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        loss += s.h_prob[4 * b];
        hit_full += s.h_prob[4 * b + 1];
        hit_train += s.h_prob[4 * b + 2];
        hit_test += s.h_prob[4 * b + 3];
    }
    Metrics m{it, owner, n ? loss / n : 0.0, n ? hit_full / n : 0.0, train ? hit_train / train : 0.0,
              n > train + val ? hit_test / (n - train - val) : 0.0};
    if (verbose) {  // the reference's log lines (gcn.h:620-632)
        printf("cross-entropy-loss = %lf\n", m.loss);
        printf("full set accuracy = %lf\n", m.acc_full);
        printf("training set accuracy = %lf\n", m.acc_train);
        printf("test set accuracy = %lf\n", m.acc_test);
    }
    return m;
}

static void ckc(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// runs `body` eagerly, or records it into `ig` (and launches the new graph), or replays `ig`
template <typename Body>
static void run_phase(SSGcnEngine::Impl& im, cudaStream_t stream, bool use_graph, SSGcnEngine::Impl::IterGraph& ig, bool online,
                      Body&& body) {
    (void)im;
    if (!use_graph) {
        body();
        return;
    }
    if (ig.exec) {
        ckc(cudaGraphLaunch(ig.exec, stream), "cudaGraphLaunch");
        im.replayed_launches += ig.launches;
        im.comm->words_sent += ig.words;
        im.comm->rounds += ig.rounds;
        if (online) im.graph_replays++;
        size_t k = 0;
        im.for_side_mats([&](DMat& m) {
            m.rows = ig.shapes[k].first;
            m.cols = ig.shapes[k].second;
            ++k;
        });
        return;
    }
    const uint64_t l0 = im.total_launches(), w0 = im.comm->words_sent, r0 = im.comm->rounds;
    cudaGraph_t g = nullptr;
    ckc(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
    g_capturing = true;
    try {
        body();
    } catch (...) {
        g_capturing = false;
        cudaStreamEndCapture(stream, &g);
        if (g) cudaGraphDestroy(g);
        throw;
    }
    g_capturing = false;
    ckc(cudaStreamEndCapture(stream, &g), "cudaStreamEndCapture");
    ckc(cudaGraphInstantiate(&ig.exec, g, 0), "cudaGraphInstantiate");
    cudaGraphDestroy(g);
    ig.launches = im.total_launches() - l0;  // counted once while recording = the launch right below
    ig.words = im.comm->words_sent - w0;
    ig.rounds = im.comm->rounds - r0;
    ig.shapes.clear();
    im.for_side_mats([&](DMat& m) { ig.shapes.push_back({m.rows, m.cols}); });
    im.graph_captures++;
    ckc(cudaGraphLaunch(ig.exec, stream), "cudaGraphLaunch");
}

void SSGcnEngine::run(uint64_t n_iters) {
    Impl& im = *impl_;
    cgb_ctx* ctx = im.main_ctx;
    cudaStream_t stream = (cudaStream_t)cgb_ctx_stream(ctx);
    for (uint64_t step = 0; step < n_iters; ++step, ++iter_) {
        const uint64_t it = iter_;
        const int ph = (int)(it % 6);
        im.comm->cur_iter = it;
        // eager in the first epoch, while tests record the transcript, or while phases are profiled; otherwise the first
        // later occurrence of this iteration index is captured and every further one replays the graph
        const bool use_graph = im.graphs_enabled && !im.profile && !im.comm->record && it >= 6 && stream != nullptr;
        auto t_deal = std::chrono::high_resolution_clock::now();
        *im.h_bias = (it - (uint64_t)ph) << 16;  // epoch part of every PRG stream id of this iteration (see stream_id)
        ck(ctx, cgb_h2d(ctx, im.d_bias, im.h_bias, 8), "h2d");
        // offline phase (dealer emulation), timed separately
        run_phase(im, stream, use_graph, im.deal_graph[ph], false, [&] { im.deal_iteration(it); });
        ck(ctx, cgb_ctx_sync(ctx), "sync");
        im.comm->barrier();  // every party has its correlations: online time below is protocol time, not dealer skew
        auto t0 = std::chrono::high_resolution_clock::now();
        seconds_offline += std::chrono::duration<double>(t0 - t_deal).count();

        const bool replay = use_graph && im.graph[ph].exec;
        if (!im.ev_on[0]) {
            ckc(cudaEventCreate(&im.ev_on[0]), "cudaEventCreate");
            ckc(cudaEventCreate(&im.ev_on[1]), "cudaEventCreate");
        }
        ckc(cudaEventRecord(im.ev_on[0], stream), "cudaEventRecord");
        run_phase(im, stream, use_graph, im.graph[ph], true, [&] { online_iteration(im, it); });
        ckc(cudaEventRecord(im.ev_on[1], stream), "cudaEventRecord");
        if (replay && ph == 1)
            for (auto& kv : im.own) im.metrics_pending.push_back(kv.first);
        ck(ctx, cgb_ctx_sync(ctx), "sync");
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, im.ev_on[0], im.ev_on[1]) == cudaSuccess) seconds_online_gpu += ms * 1e-3;
        }
        for (int owner : im.metrics_pending) metrics_.push_back(owner_metrics(im, it, owner, verbose));
        im.metrics_pending.clear();
        const double dt = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        seconds_online += dt;
        if (verbose) printf("::iteration took %lf seconds\n", dt);
    }
}

double SSGcnEngine::seconds_residual_host() const { return 0.0; }  // the stand-ins run on the device since round 1b
uint64_t SSGcnEngine::replayed_launches() const { return impl_->replayed_launches; }
uint64_t SSGcnEngine::eager_launches() const { return impl_->total_launches(); }
uint64_t SSGcnEngine::graph_replays() const { return impl_->graph_replays; }

std::vector<uint64_t> SSGcnEngine::download(int owner, int role, const std::string& name, uint32_t* rows, uint32_t* cols) {
    Impl& im = *impl_;
    Side* s = im.side(owner, role);
    if (!s) throw std::runtime_error("download: side not hosted here");
    DMat* m = nullptr;
    if (name == "X") m = &s->X;
    else if (name == "W0") m = &s->W[0];
    else if (name == "W1") m = &s->W[1];
    else if (name == "z0") m = &s->z[0];
    else if (name == "z1") m = &s->z[1];
    else if (name == "h_t0") m = &s->h_t[0];
    else if (name == "h_t1") m = &s->h_t[1];
    else if (name == "g") m = &s->g;
    else if (name == "V") m = &s->V;
    else if (name == "Xp") m = &s->Xp;
    else throw std::runtime_error("download: unknown tensor " + name);
    std::vector<uint64_t> out(m->n());
    if (!out.empty()) ck(im.ctx, cgb_d2h(im.ctx, out.data(), m->p, out.size() * 8), "d2h");
    ck(im.ctx, cgb_ctx_sync(im.ctx), "sync");
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    return out;
}

}  // namespace cognn
