// cognn_shim.h -- header-only C++17 binding of the reference's "level B" primitive API onto the C ABI of
// include/cognn_b200.h.  It re-creates the NAMES AND SIGNATURES the CoGNN operators call but /root/reference does
// not define (they live in the absent Task-Worker tree: SCIHarness.h, ObliviousMapper.h, SecureAggregation.h,
// TaskUtil.h), so that algo_kernels/vertex_centric/optimize-gcn/gcn.h-shaped operator code compiles against it:
//
//   typedefs ShareVec / ShareTensor / ShareVecVec / DoubleTensor / PosVec, transpose(), toShareVec()   task.h:237-272
//   GNNParam::getGNNParam().readConfig()                                                               task.h:78-170
//   sci::ALICE / sci::BOB, sci::twoPartyGCNMatMul            gcn.h:233,665,671,710
//   sci::twoPartyGCNVectorScale                              gcn.h:247,476
//   sci::twoPartyGCNMatrixScale / twoPartyGCNApplyGradient   gcn.h:676,678,723,730,764
//   sci::twoPartyGCNCondVectorAddition                       gcn.h:456
//   sci::getPlainShareVecVec                                 gcn.h:604
//   prefix_network_aggregate(..., AggregationOp::ADD_AGG, ...)          gcn.h:328
//   client_oblivious_mapper_online / server_oblivious_mapper_online     ssk.h:752,760,818,848 / 1011,1016,1057,1075
//   CryptoUtil::intoShares / encodeDoubleAsFixedPoint / mergeShareAsDouble   gcn.h:70,80,96,220
//
// Calling convention kept from the reference: caller-owned STL containers by reference, the callee resizes outputs,
// in/out aliasing is allowed (results are computed out of place), errors are fatal (printf + exit(-1), ssk.h:794-797).
// The two parties of a call run the same sequence of primitives (as ALICE and BOB threads do in ssk.h:702-704 /
// 926-928); a per-owner operation counter keeps their dealer PRG streams aligned.
//
// NOT provided (2PC-RESIDUAL, stays on the reference's SCI/OT backend): sci::twoPartyGCNRelu,
// sci::twoPartyGCNForwardNNPredictionWithoutWeight, sci::twoPartyGCNBackwardNNWithoutAH.
// This is the API-faithful path: every call moves host vectors to the GPU and back.  The device-resident fast path is
// the engine (engine.h); both produce the same reconstructed values.
#pragma once
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/cognn_b200.h"

#ifndef SCALER_BIT_LENGTH
#define SCALER_BIT_LENGTH CGB_SCALER_BITS
#endif

#ifndef TASK_H_
// ---- stand-alone mode: the vocabulary of the reference's include/task/task.h:78-272 (GNNParam, share typedefs, PosVec) --------
typedef std::vector<uint64_t> ShareVec;
typedef std::vector<std::vector<double>> DoubleTensor;
typedef std::vector<std::vector<uint64_t>> ShareTensor;
typedef std::vector<ShareTensor> ShareTensorVec;
typedef std::vector<ShareVec> ShareVecVec;
struct PosVec {
    std::vector<uint64_t> pos;
};

inline ShareTensor transpose(const ShareTensor& st) {
    if (st.empty()) return ShareTensor();
    ShareTensor out(st[0].size(), ShareVec(st.size()));
    for (size_t i = 0; i < st.size(); ++i)
        for (size_t j = 0; j < st[i].size(); ++j) out[j][i] = st[i][j];
    return out;
}
inline ShareVec toShareVec(int hotIndex, int vecSize) {  // one-hot label in fixed point (gcn.h:575)
    ShareVec v(vecSize, 0);
    if (hotIndex >= 0 && hotIndex < vecSize) v[hotIndex] = 1ull << SCALER_BIT_LENGTH;
    return v;
}

class GNNParam {  // task.h:78-170
    GNNParam() {}

public:
    int num_layers = 2, num_labels = 0, input_dim = 0, hidden_dim = 0, num_samples = 0, num_edges = 0;
    double learning_rate = 0, train_ratio = 0, val_ratio = 0, test_ratio = 0;
    static GNNParam& getGNNParam() {
        static GNNParam instance;
        return instance;
    }
    void readConfig(const std::string& file_name) {
        std::ifstream fin(file_name);
        if (!fin.is_open()) {
            std::cerr << "Failed to open the file: " << file_name << std::endl;
            return;
        }
        std::string param;
        char colon;
        while (fin >> param >> colon) {
            if (colon != ':') {
                std::cerr << "Invalid format: expected a colon after " << param << std::endl;
                break;
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
                std::cerr << "Unknown parameter: " << param << std::endl;
                break;
            }
        }
    }
};

#else
// ---- the reference's own task.h is in the translation unit (its operator headers are being compiled against this shim): it
// declares these helpers and leaves their definitions to Task-Worker (absent); define them here, once, inline
inline ShareTensor transpose(const ShareTensor& st) {
    if (st.empty()) return ShareTensor();
    ShareTensor out(st[0].size(), ShareVec(st.size()));
    for (size_t i = 0; i < st.size(); ++i)
        for (size_t j = 0; j < st[i].size(); ++j) out[j][i] = st[i][j];
    return out;
}
inline ShareTensor toShareTensor(const ShareVec& sv) { return ShareTensor(1, sv); }
inline ShareVec toShareVec(const ShareTensor& st) {
    ShareVec v;
    for (const auto& r : st) v.insert(v.end(), r.begin(), r.end());
    return v;
}
#ifndef SCALER_BIT_LENGTH
#define SCALER_BIT_LENGTH CGB_SCALER_BITS
#endif
inline ShareVec toShareVec(int hotIndex, int vecSize) {  // one-hot label in fixed point (gcn.h:575)
    ShareVec v(vecSize, 0);
    if (hotIndex >= 0 && hotIndex < vecSize) v[hotIndex] = 1ull << SCALER_BIT_LENGTH;
    return v;
}
#endif  // TASK_H_

enum class AggregationOp { ADD_AGG, MIN_AGG, MAX_AGG };

namespace cognn_shim {

[[noreturn]] inline void fatal(const char* what, const char* detail) {
    printf("%s: %s\n", what, detail ? detail : "");  // the reference's error behaviour (ssk.h:794-797)
    exit(-1);
}

// message plane between the two threads / processes of one (owner, helper) pair
struct Channel {
    virtual ~Channel() {}
    virtual void send(const std::vector<uint64_t>& v) = 0;
    virtual void recv(std::vector<uint64_t>& v) = 0;  // whole message
};

// both ends in one process (ALICE and BOB threads, like the reference's single-host runs with -c 0)
class InProcPipe {
public:
    struct End : Channel {
        InProcPipe* p;
        int side;
        void send(const std::vector<uint64_t>& v) override {
            std::lock_guard<std::mutex> l(p->m);
            p->q[side].push_back(v);
            p->cv.notify_all();
        }
        void recv(std::vector<uint64_t>& v) override {
            std::unique_lock<std::mutex> l(p->m);
            p->cv.wait(l, [&] { return !p->q[1 - side].empty(); });
            v = std::move(p->q[1 - side].front());
            p->q[1 - side].pop_front();
        }
    };
    InProcPipe() {
        a.p = b.p = this;
        a.side = 0;
        b.side = 1;
    }
    End a, b;

private:
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::vector<uint64_t>> q[2];
};

// dealer PRG stream ids: same layout as the engine (DESIGN.md "Randomness"); `op` plays the role of the iteration
enum Kind : uint64_t { K_MM_U0 = 5, K_MM_U1, K_MM_V0, K_MM_V1, K_MM_Z0, K_RM_A0, K_RM_A1, K_RM_B0, K_RM_B1, K_RM_C0,
                       K_OM_R = 32, K_OM_S = 33, K_SPLIT = 34, K_RESHARE = 35 };
inline uint64_t stream_id(uint64_t kind, uint64_t op, uint64_t owner, uint64_t sub) {
    return (kind << 48) | (op << 16) | (owner << 8) | sub;
}
// dealer streams of an (owner, helper) pair: 4 bits of sub-stream, 4 bits of helper id in the low byte (T <= 16)
inline uint64_t pair_sub(uint64_t helper, uint64_t sub) { return ((helper & 15) << 4) | (sub & 15); }

class Runtime {
public:
    Runtime(int tileIndex, int tileNum, int device, const uint32_t key_[8]) : tileIndex(tileIndex), tileNum(tileNum), device(device) {
        for (int i = 0; i < 8; ++i) key[i] = key_[i];
    }
    ~Runtime() {
        for (auto& kv : ctxs) cgb_ctx_destroy(kv.second);
    }
    // the thread that will issue primitives for (coTid, party) registers its channel to the matching thread of coTid
    void connect(uint64_t coTid, int party, Channel* ch) {
        std::lock_guard<std::mutex> l(mu);
        chans[{coTid, party}] = ch;
    }
    void connect_owned(uint64_t coTid, int party, Channel* ch) {  // the runtime deletes the channel
        connect(coTid, party, ch);
        std::lock_guard<std::mutex> l(mu);
        owned.emplace_back(ch);
    }
    static Runtime*& current_ptr() {
        static thread_local Runtime* rt = nullptr;
        return rt;
    }
    static void bind_thread(Runtime* rt) { current_ptr() = rt; }
    // one party per process (the reference's deployment: its engine spawns the ALICE / BOB threads itself, so nothing can bind
    // them): the runtime every unbound thread falls back to
    static Runtime*& process_ptr() {
        static Runtime* rt = nullptr;
        return rt;
    }
    static void bind_process(Runtime* rt) { process_ptr() = rt; }
    static Runtime& current() {
        if (current_ptr()) return *current_ptr();
        if (process_ptr()) return *process_ptr();
        fatal("cognn_shim", "no Runtime bound to this thread or process (Runtime::bind_thread / bind_process)");
    }
    cgb_ctx* ctx(uint64_t coTid, int party) {
        std::lock_guard<std::mutex> l(mu);
        auto key_ = std::make_pair(coTid, party);
        auto it = ctxs.find(key_);
        if (it != ctxs.end()) return it->second;
        cgb_ctx* c = nullptr;
        if (cgb_ctx_create(device, &c) != CGB_OK) fatal("cgb_ctx_create", cgb_last_error(nullptr));
        ctxs[key_] = c;
        return c;
    }
    Channel& chan(uint64_t coTid, int party) {
        std::lock_guard<std::mutex> l(mu);
        auto it = chans.find({coTid, party});
        if (it == chans.end()) fatal("cognn_shim", "no channel connected for this (coTid, party)");
        return *it->second;
    }
    // one counter per (owner, helper) pair and share: the ALICE thread of the owner and the BOB thread of the helper issue the
    // same sequence of two-party primitives (ssk.h:702-704 / 926-928), other pairs run concurrently with their own sequence
    uint64_t next_op(uint64_t owner, uint64_t helper, int share) {
        std::lock_guard<std::mutex> l(mu);
        return ops[{owner * 256 + helper, share}]++;
    }
    int tileIndex, tileNum, device;
    uint32_t key[8];

private:
    std::mutex mu;
    std::map<std::pair<uint64_t, int>, cgb_ctx*> ctxs;
    std::map<std::pair<uint64_t, int>, Channel*> chans;
    std::map<std::pair<uint64_t, int>, uint64_t> ops;
    std::vector<std::unique_ptr<Channel>> owned;
};

// ---- small RAII device buffer + host <-> device helpers -------------------------------------------------------------
struct Dev {
    cgb_ctx* c;
    uint64_t* p = nullptr;
    size_t n = 0;
    Dev(cgb_ctx* c, size_t n) : c(c), n(n) {
        void* q = nullptr;
        if (cgb_malloc(c, (n ? n : 1) * 8, &q) != CGB_OK) fatal("cgb_malloc", cgb_last_error(c));
        p = (uint64_t*)q;
    }
    Dev(const Dev&) = delete;
    ~Dev() { cgb_free(c, p); }
    void up(const std::vector<uint64_t>& h) {
        if (h.size() != n) fatal("cognn_shim", "upload size mismatch");
        if (n && cgb_h2d(c, p, h.data(), n * 8) != CGB_OK) fatal("cgb_h2d", cgb_last_error(c));
        cgb_ctx_sync(c);
    }
    std::vector<uint64_t> down() const {
        std::vector<uint64_t> h(n);
        if (n && cgb_d2h(c, h.data(), p, n * 8) != CGB_OK) fatal("cgb_d2h", cgb_last_error(c));
        cgb_ctx_sync(c);
        return h;
    }
};
inline void ok(cgb_ctx* c, int rc, const char* what) {
    if (rc != CGB_OK) fatal(what, cgb_last_error(c));
}
inline std::vector<uint64_t> flatten(const ShareVecVec& m, size_t* rows, size_t* cols) {
    *rows = m.size();
    *cols = m.empty() ? 0 : m[0].size();
    std::vector<uint64_t> f(*rows * *cols);
    for (size_t i = 0; i < *rows; ++i) {
        if (m[i].size() != *cols) fatal("cognn_shim", "ragged share matrix");
        for (size_t j = 0; j < *cols; ++j) f[i * *cols + j] = m[i][j];
    }
    return f;
}
inline void unflatten(const std::vector<uint64_t>& f, size_t rows, size_t cols, ShareVecVec& out) {
    ShareVecVec r(rows, ShareVec(cols));
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < cols; ++j) r[i][j] = f[i * cols + j];
    out.swap(r);  // out may alias an input: assign only at the end
}
inline void prg(cgb_ctx* c, const uint32_t key[8], uint64_t sid, Dev& d) { ok(c, cgb_prg_fill(c, key, sid, 0, d.p, d.n), "cgb_prg_fill"); }
// exchange: send mine, receive the peer's message of the same size, return it on the device
inline void swap_msgs(Channel& ch, const Dev& mine, Dev& peer) {
    ch.send(mine.down());
    std::vector<uint64_t> v;
    ch.recv(v);
    peer.up(v);
}
struct Call {  // common prologue of every two-party primitive
    Runtime& rt;
    int share;
    uint64_t owner, helper, op;
    cgb_ctx* c;
    Channel& ch;
    Call(uint64_t coTid, int party)
        : rt(Runtime::current()), share(party == 1 ? 0 : 1), owner(party == 1 ? (uint64_t)rt.tileIndex : coTid),
          helper(party == 1 ? coTid : (uint64_t)rt.tileIndex), op(rt.next_op(owner, helper, share)), c(rt.ctx(coTid, party)),
          ch(rt.chan(coTid, party)) {
        if (party != 1 && party != 2) fatal("cognn_shim", "party must be sci::ALICE (1) or sci::BOB (2)");
    }
    uint64_t sid(uint64_t kind, uint64_t sub = 0) const { return stream_id(kind, op, owner, pair_sub(helper, sub)); }
};

}  // namespace cognn_shim

// ---------------------------------------------------------------------------------------------------------------------
class CryptoUtil {
public:
    // the reference's main() configures a singleton (harness.cpp:119-121); the Paillier / FHE set-up it asks for belongs to the
    // HE task queue that the secret-shared GCN path does not use ("not used", README.md:107)
    static CryptoUtil& getInstance() {
        static CryptoUtil inst;
        return inst;
    }
    void tileIndexIs(size_t t) { tileIndex_ = t; }
    size_t tileIndex() const { return tileIndex_; }
    void setUpPaillierCipher() {}
    void setUpFHECipher() {}

    static uint64_t encodeDoubleAsFixedPoint(double x) { return (uint64_t)(int64_t)(x * (double)(1ull << SCALER_BIT_LENGTH)); }
    static double decodeFixedPointAsDouble(uint64_t v) { return (double)(int64_t)v / (double)(1ull << SCALER_BIT_LENGTH); }
    // s1 = next word of the party's split stream, s0 = enc(x) - s1 (DESIGN.md "Frozen semantics").  The reference calls this
    // from OpenMP workers (optimize-gcn/gcn.h:86-99 intoShareTensor): ONE process-wide pool under a mutex, refilled from
    // consecutive PRG blocks of one stream, so no two calls -- on whatever thread -- ever draw the same mask word.
    static void intoShares(double x, uint64_t& s0, uint64_t& s1) {
        static std::mutex mu;
        static std::vector<uint64_t> pool;
        static size_t pos = 0;
        static uint64_t refill = 0;
        std::lock_guard<std::mutex> l(mu);
        if (pos == pool.size()) {
            auto& rt = cognn_shim::Runtime::current();
            cgb_ctx* c = rt.ctx((uint64_t)-1, 1);
            cognn_shim::Dev d(c, 1 << 16);
            cognn_shim::ok(c, cgb_prg_fill(c, rt.key, cognn_shim::stream_id(cognn_shim::K_SPLIT, refill, rt.tileIndex, 0),
                                           0, d.p, d.n), "cgb_prg_fill");
            pool = d.down();
            pos = 0;
            ++refill;
        }
        s1 = pool[pos++];
        s0 = encodeDoubleAsFixedPoint(x) - s1;
    }
    static double mergeShareAsDouble(uint64_t s0, uint64_t s1) { return decodeFixedPointAsDouble(s0 + s1); }

private:
    size_t tileIndex_ = 0;
};

namespace sci {

const int ALICE = 1;
const int BOB = 2;

// ---- plaintext helpers the operators use for the FedAvg sums and for the metrics the owner prints (gcn.h:603-632, 753-777) ----
inline void plaintext_add_matrix_in_place(ShareVecVec& a, const ShareVecVec& b) {
    if (a.size() != b.size()) cognn_shim::fatal("plaintext_add_matrix_in_place", "row count mismatch");
    for (size_t i = 0; i < a.size(); ++i) {
        if (a[i].size() != b[i].size()) cognn_shim::fatal("plaintext_add_matrix_in_place", "column count mismatch");
        for (size_t j = 0; j < a[i].size(); ++j) a[i][j] += b[i][j];
    }
}
inline ShareVecVec plaintext_add_matrix(const ShareVecVec& a, const ShareVecVec& b) {
    ShareVecVec r = a;
    plaintext_add_matrix_in_place(r, b);
    return r;
}
inline double cross_entropy_loss(const DoubleTensor& y, const DoubleTensor& p) {  // mean over rows of -log p[label]
    if (y.empty()) return 0.0;
    double tot = 0.0;
    for (size_t i = 0; i < y.size(); ++i)
        for (size_t j = 0; j < y[i].size(); ++j)
            if (y[i][j] != 0.0) tot -= y[i][j] * std::log(p[i][j] > 1e-30 ? p[i][j] : 1e-30);
    return tot / (double)y.size();
}
inline size_t argmax_row(const std::vector<double>& r) {
    size_t b = 0;
    for (size_t j = 1; j < r.size(); ++j)
        if (r[j] > r[b]) b = j;
    return b;
}
inline double accuracy(const DoubleTensor& y, const DoubleTensor& p) {
    if (y.empty()) return 0.0;
    size_t hit = 0;
    for (size_t i = 0; i < y.size(); ++i) hit += argmax_row(y[i]) == argmax_row(p[i]);
    return (double)hit / (double)y.size();
}
inline double accuracy(const DoubleTensor& y, const DoubleTensor& p, const std::vector<bool>& mask) {  // rows with mask only
    size_t hit = 0, n = 0;
    for (size_t i = 0; i < y.size(); ++i)
        if (mask[i]) {
            ++n;
            hit += argmax_row(y[i]) == argmax_row(p[i]);
        }
    return n ? (double)hit / (double)n : 0.0;
}
inline size_t count_true(const std::vector<bool>& v) {
    size_t n = 0;
    for (bool b : v) n += b;
    return n;
}
template <typename T>
inline void print_vector(const std::vector<T>& v, size_t limit = (size_t)-1) {
    for (size_t i = 0; i < v.size() && i < limit; ++i) std::cout << v[i] << " ";
    std::cout << std::endl;
}
template <typename T>
inline void print_vector_of_vector(const std::vector<std::vector<T>>& m, size_t limit = (size_t)-1) {
    for (size_t i = 0; i < m.size() && i < limit; ++i) print_vector(m[i]);
}
// GCN_LOG debugging aid of the reference: opens the shares towards ALICE and prints the fixed-point values (insecure by nature;
// only compiled into debugging builds of the operator headers, -DGCN_LOG)
inline void printShareVecVec(const ShareVecVec& m, uint64_t coTid, int party) {
    auto& rt = cognn_shim::Runtime::current();
    cognn_shim::Channel& ch = rt.chan(coTid, party);
    size_t rows, cols;
    std::vector<uint64_t> x = cognn_shim::flatten(m, &rows, &cols);
    if (party == 2) {
        ch.send(x);
        return;
    }
    std::vector<uint64_t> other;
    ch.recv(other);
    static std::mutex mu;
    static std::map<uint64_t, int> calls;
    std::lock_guard<std::mutex> l(mu);
    const int call = calls[coTid]++;
    std::string out;
    char buf[64];
    for (size_t i = 0; i < rows; ++i) {
        snprintf(buf, sizeof buf, "\n[svv peer=%llu call=%d row=%zu]", (unsigned long long)coTid, call, i);
        out += buf;
        for (size_t j = 0; j < cols; ++j) {
            snprintf(buf, sizeof buf, " %.6f",
                     (double)(int64_t)(x[i * cols + j] + (i * cols + j < other.size() ? other[i * cols + j] : 0)) / (double)(1ull << SCALER_BIT_LENGTH));
            out += buf;
        }
    }
    out += "\n";
    fputs(out.c_str(), stdout);
    fflush(stdout);
}

// C = trunc(A * B) on shares (Beaver triple from the dealer PRG; one message each way)
inline void twoPartyGCNMatMul(const ShareVecVec& A, const ShareTensor& B, ShareVecVec& C, uint64_t coTid, int party) {
    using namespace cognn_shim;
    Call k(coTid, party);
    size_t M, K, K2, N;
    std::vector<uint64_t> a = flatten(A, &M, &K), b = flatten(B, &K2, &N);
    if (K != K2) fatal("twoPartyGCNMatMul", "inner dimensions differ");
    cgb_ctx* c = k.c;
    Dev dA(c, M * K), dB(c, K * N), U(c, M * K), V(c, K * N), Z(c, M * N), mine(c, M * K + K * N), peer(c, M * K + K * N), out(c, M * N);
    dA.up(a);
    dB.up(b);
    if (k.share == 0) {
        prg(c, k.rt.key, k.sid(K_MM_U0), U);
        prg(c, k.rt.key, k.sid(K_MM_V0), V);
        prg(c, k.rt.key, k.sid(K_MM_Z0), Z);
    } else {  // dealer emulation: Z1 = (U0+U1)(V0+V1) - Z0
        Dev U0(c, M * K), V0(c, K * N), Z0(c, M * N);
        prg(c, k.rt.key, k.sid(K_MM_U0), U0);
        prg(c, k.rt.key, k.sid(K_MM_V0), V0);
        prg(c, k.rt.key, k.sid(K_MM_Z0), Z0);
        prg(c, k.rt.key, k.sid(K_MM_U1), U);
        prg(c, k.rt.key, k.sid(K_MM_V1), V);
        ok(c, cgb_add(c, U0.p, U.p, U0.p, U0.n), "cgb_add");
        ok(c, cgb_add(c, V0.p, V.p, V0.p, V0.n), "cgb_add");
        ok(c, cgb_matmul(c, U0.p, V0.p, Z.p, M, K, N, 0, 0), "cgb_matmul");
        ok(c, cgb_sub(c, Z.p, Z0.p, Z.p, Z.n), "cgb_sub");
        cgb_ctx_sync(c);
    }
    ok(c, cgb_sub(c, dA.p, U.p, mine.p, M * K), "cgb_sub");
    ok(c, cgb_sub(c, dB.p, V.p, mine.p + M * K, K * N), "cgb_sub");
    swap_msgs(k.ch, mine, peer);
    ok(c, cgb_add(c, mine.p, peer.p, mine.p, mine.n), "cgb_add");
    ok(c, cgb_beaver_matmul_finish(c, mine.p, mine.p + M * K, U.p, V.p, Z.p, out.p, M, K, N, k.share, SCALER_BIT_LENGTH),
       "cgb_beaver_matmul_finish");
    unflatten(out.down(), M, N, C);
}

namespace detail {
// out = [trunc](x * s[row]) with s = s_ALICE + s_BOB: in CoGNN-Opt the scaler is private to ALICE and BOB passes zeros
// (ssk.h:985-994); in original-gcn BOB supplies the destination normaliser of mirror edges instead (ssk.h:1043)
inline void rowmul(const ShareVecVec& in, const std::vector<uint64_t>& scaler, ShareVecVec& out, uint64_t coTid, int party, int f) {
    using namespace cognn_shim;
    Call k(coTid, party);
    size_t rows, D;
    std::vector<uint64_t> x = flatten(in, &rows, &D);
    if (scaler.size() != rows) fatal("twoPartyGCNVectorScale", "one scaler per row expected");
    cgb_ctx* c = k.c;
    Dev dx(c, rows * D), a(c, rows * D), b(c, rows), cc(c, rows * D), mine(c, rows * D + rows), peer(c, rows * D + rows), res(c, rows * D);
    dx.up(x);
    if (k.share == 0) {
        prg(c, k.rt.key, k.sid(K_RM_A0), a);
        prg(c, k.rt.key, k.sid(K_RM_B0), b);
        prg(c, k.rt.key, k.sid(K_RM_C0), cc);
    } else {
        Dev a0(c, rows * D), b0(c, rows), c0(c, rows * D), zm(c, rows * D), zv(c, rows);
        prg(c, k.rt.key, k.sid(K_RM_A0), a0);
        prg(c, k.rt.key, k.sid(K_RM_B0), b0);
        prg(c, k.rt.key, k.sid(K_RM_C0), c0);
        prg(c, k.rt.key, k.sid(K_RM_A1), a);
        prg(c, k.rt.key, k.sid(K_RM_B1), b);
        ok(c, cgb_add(c, a0.p, a.p, a0.p, a0.n), "cgb_add");
        ok(c, cgb_add(c, b0.p, b.p, b0.p, b0.n), "cgb_add");
        ok(c, cgb_memset(c, zm.p, 0, zm.n * 8), "memset");
        ok(c, cgb_memset(c, zv.p, 0, zv.n * 8), "memset");
        ok(c, cgb_sub(c, zm.p, c0.p, c0.p, c0.n), "cgb_sub");  // -c0
        ok(c, cgb_rowmul_beaver_finish(c, a0.p, b0.p, zm.p, zv.p, c0.p, cc.p, rows, D, 0, -1), "rowmul(dealer)");
        cgb_ctx_sync(c);
    }
    ok(c, cgb_sub(c, dx.p, a.p, mine.p, rows * D), "cgb_sub");
    if (k.share == 0) {
        Dev s(c, rows);
        s.up(scaler);
        ok(c, cgb_sub(c, s.p, b.p, mine.p + rows * D, rows), "cgb_sub");
        cgb_ctx_sync(c);
    } else {
        Dev s(c, rows);
        s.up(scaler);
        ok(c, cgb_sub(c, s.p, b.p, mine.p + rows * D, rows), "cgb_sub");
        cgb_ctx_sync(c);
    }
    swap_msgs(k.ch, mine, peer);
    ok(c, cgb_add(c, mine.p, peer.p, mine.p, mine.n), "cgb_add");
    ok(c, cgb_rowmul_beaver_finish(c, mine.p, mine.p + rows * D, a.p, b.p, cc.p, res.p, rows, D, k.share, f), "cgb_rowmul_beaver_finish");
    unflatten(res.down(), rows, D, out);
}
}  // namespace detail

inline void twoPartyGCNVectorScale(const ShareVecVec& in, const std::vector<uint64_t>& scaler, ShareVecVec& out, bool /*isSigned*/,
                                   uint64_t coTid, int party) {
    detail::rowmul(in, scaler, out, coTid, party, SCALER_BIT_LENGTH);
}

// out = v + (cond ? u : 0); cond is private to ALICE (BOB passes all-true, ssk.h:1124-1126): MUX via a Beaver product
inline void twoPartyGCNCondVectorAddition(const ShareVecVec& v, const ShareVecVec& u, const std::vector<bool>& cond, ShareVecVec& out,
                                          uint64_t coTid, int party) {
    // the selector is ALICE's alone: BOB's argument carries no information (all true, ssk.h:1124-1126) and must not be added in
    std::vector<uint64_t> sel(cond.size(), 0);
    if (party == ALICE)
        for (size_t i = 0; i < cond.size(); ++i) sel[i] = cond[i] ? 1 : 0;
    ShareVecVec gated;
    detail::rowmul(u, sel, gated, coTid, party, -1);
    ShareVecVec r(v.size());
    for (size_t i = 0; i < v.size(); ++i) {
        r[i].resize(v[i].size());
        for (size_t j = 0; j < v[i].size(); ++j) r[i][j] = v[i][j] + gated[i][j];
    }
    out.swap(r);
}

// public scalar: share-local (both parties pass the same value; gcn.h:676,764)
inline void twoPartyGCNMatrixScale(const ShareTensor& in, uint64_t scaler, ShareTensor& out, uint64_t coTid, int party) {
    using namespace cognn_shim;
    auto& rt = Runtime::current();
    cgb_ctx* c = rt.ctx(coTid, party);
    size_t rows, cols;
    std::vector<uint64_t> x = flatten(in, &rows, &cols);
    Dev d(c, x.size());
    d.up(x);
    ok(c, cgb_scale_public(c, d.p, scaler, d.p, d.n, SCALER_BIT_LENGTH, party == ALICE ? 0 : 1), "cgb_scale_public");
    unflatten(d.down(), rows, cols, out);
}
inline void twoPartyGCNApplyGradient(const ShareTensor& W, const ShareTensor& dW, uint64_t lr, ShareTensor& out, uint64_t coTid, int party) {
    using namespace cognn_shim;
    auto& rt = Runtime::current();
    cgb_ctx* c = rt.ctx(coTid, party);
    size_t rows, cols, r2, c2;
    std::vector<uint64_t> w = flatten(W, &rows, &cols), g = flatten(dW, &r2, &c2);
    if (rows != r2 || cols != c2) fatal("twoPartyGCNApplyGradient", "shape mismatch");
    Dev dw(c, w.size()), dg(c, g.size());
    dw.up(w);
    dg.up(g);
    ok(c, cgb_apply_gradient(c, dw.p, dg.p, lr, dw.p, dw.n, SCALER_BIT_LENGTH, party == ALICE ? 0 : 1), "cgb_apply_gradient");
    unflatten(dw.down(), rows, cols, out);
}
// opening towards ALICE (gcn.h:604: the owner learns the predictions); BOB gets an empty tensor
inline void getPlainShareVecVec(const ShareTensor& st, DoubleTensor& plain, uint64_t coTid, int party) {
    using namespace cognn_shim;
    auto& rt = Runtime::current();
    Channel& ch = rt.chan(coTid, party);
    size_t rows, cols;
    std::vector<uint64_t> x = flatten(st, &rows, &cols);
    if (party == BOB) {
        ch.send(x);
        plain.clear();
        return;
    }
    std::vector<uint64_t> other;
    ch.recv(other);
    if (other.size() != x.size()) fatal("getPlainShareVecVec", "share size mismatch");
    cgb_ctx* c = rt.ctx(coTid, party);
    Dev a(c, x.size()), b(c, x.size());
    a.up(x);
    b.up(other);
    void* dv = nullptr;
    ok(c, cgb_malloc(c, (x.size() ? x.size() : 1) * 8, &dv), "cgb_malloc");
    ok(c, cgb_open_decode(c, a.p, b.p, (double*)dv, x.size(), SCALER_BIT_LENGTH), "cgb_open_decode");
    std::vector<double> h(x.size());
    if (!h.empty()) ok(c, cgb_d2h(c, h.data(), dv, h.size() * 8), "cgb_d2h");
    cgb_ctx_sync(c);
    cgb_free(c, dv);
    plain.assign(rows, std::vector<double>(cols));
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < cols; ++j) plain[i][j] = h[i * cols + j];
}

#ifdef COGNN_SHIM_IDEAL_NONLINEAR
// ---- 2PC-RESIDUAL stand-ins: IDEAL FUNCTIONALITY, **NOT SECURE** -------------------------------------------------------------
// sci::twoPartyGCNRelu (gcn.h:549), sci::twoPartyGCNForwardNNPredictionWithoutWeight (gcn.h:578,591) and
// sci::twoPartyGCNBackwardNNWithoutAH (gcn.h:705) are comparison / exponentiation protocols of the reference's MPC backend
// (SCI-SilentOT) and are NOT replaced by this library.  Only when COGNN_SHIM_IDEAL_NONLINEAR is defined -- to run an epoch
// end to end on test data -- the shim defines them as the same stand-in the engine uses: BOB sends its share to ALICE, ALICE
// evaluates the function on the RECONSTRUCTED values (cgb_ideal_*: she sees the plaintext) and re-shares the result with a
// dealer PRG stream both sides can derive.  Never define the macro in a deployment.
namespace detail {
// BOB: send shares, take the dealer stream(s) as new share(s).  ALICE: receive, return the peer's words.
inline bool ideal_exchange(cognn_shim::Call& k, int party, const std::vector<uint64_t>& mine, std::vector<uint64_t>& peer) {
    if (party == BOB) {
        k.ch.send(mine);
        return false;
    }
    k.ch.recv(peer);
    if (peer.size() != mine.size()) cognn_shim::fatal("cognn_shim ideal stand-in", "share size mismatch");
    return true;
}
inline void ideal_reshare(cognn_shim::Call& k, int party, int sub, const cognn_shim::Dev* plain, size_t rows, size_t cols, ShareVecVec& out) {
    using namespace cognn_shim;
    Dev r(k.c, rows * cols);
    prg(k.c, k.rt.key, k.sid(K_RESHARE, sub), r);
    if (party == BOB) {
        unflatten(r.down(), rows, cols, out);
        return;
    }
    Dev o(k.c, rows * cols);
    ok(k.c, cgb_sub(k.c, plain->p, r.p, o.p, o.n), "cgb_sub");
    unflatten(o.down(), rows, cols, out);
}
}  // namespace detail

inline void twoPartyGCNRelu(const ShareVecVec& in, ShareTensor& out, uint64_t coTid, int party) {
    using namespace cognn_shim;
    Call k(coTid, party);
    size_t rows, cols;
    std::vector<uint64_t> x = flatten(in, &rows, &cols), other;
    if (!detail::ideal_exchange(k, party, x, other)) return detail::ideal_reshare(k, party, 0, nullptr, rows, cols, out);
    Dev a(k.c, x.size()), b(k.c, x.size()), plain(k.c, x.size());
    a.up(x);
    b.up(other);
    ok(k.c, cgb_ideal_relu(k.c, a.p, b.p, plain.p, x.size()), "cgb_ideal_relu");
    detail::ideal_reshare(k, party, 0, &plain, rows, cols, out);
}

// p = softmax(z) (fixed point), p_minus_y = p - label; ALICE passes the one-hot label rows, BOB zeros (gcn.h:574-590);
// the rows outside the training set are zeroed by the caller afterwards (gcn.h:639-641)
inline void twoPartyGCNForwardNNPredictionWithoutWeight(const ShareVecVec& z, const ShareVecVec& label, ShareVecVec& p,
                                                        ShareVecVec& p_minus_y, uint64_t coTid, int party) {
    using namespace cognn_shim;
    Call k(coTid, party);
    size_t rows, cols, lr, lc;
    std::vector<uint64_t> x = flatten(z, &rows, &cols), lab = flatten(label, &lr, &lc), other, other_lab;
    if (lr != rows || lc != cols) fatal("twoPartyGCNForwardNNPredictionWithoutWeight", "label shape differs from the logits");
    const bool alice = detail::ideal_exchange(k, party, x, other);
    if (!alice) {
        k.ch.send(lab);
        detail::ideal_reshare(k, party, 0, nullptr, rows, cols, p);
        return detail::ideal_reshare(k, party, 1, nullptr, rows, cols, p_minus_y);
    }
    k.ch.recv(other_lab);
    if (other_lab.size() != lab.size()) fatal("twoPartyGCNForwardNNPredictionWithoutWeight", "label size mismatch");
    std::vector<int32_t> cls(rows, 0);  // class of each row = position of the 1.0 in the reconstructed one-hot label
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < cols; ++j)
            if (lab[i * cols + j] + other_lab[i * cols + j] != 0) cls[i] = (int32_t)j;
    Dev a(k.c, x.size()), b(k.c, x.size()), P(k.c, x.size()), D(k.c, x.size());
    a.up(x);
    b.up(other);
    void* dl = nullptr;
    ok(k.c, cgb_malloc(k.c, (rows ? rows : 1) * sizeof(int32_t), &dl), "cgb_malloc");
    if (rows) ok(k.c, cgb_h2d(k.c, dl, cls.data(), rows * sizeof(int32_t)), "cgb_h2d");
    ok(k.c, cgb_ideal_softmax(k.c, a.p, b.p, (const int32_t*)dl, rows, (uint32_t)cols, rows, SCALER_BIT_LENGTH, P.p, D.p),
       "cgb_ideal_softmax");
    cgb_ctx_sync(k.c);
    cgb_free(k.c, dl);
    // rows whose reconstructed label is all zero (inference: zero_label on both sides) keep p_minus_y = p - 0
    bool any_unlabelled = false;
    std::vector<uint8_t> has(rows, 0);
    for (size_t i = 0; i < rows; ++i) {
        for (size_t j = 0; j < cols; ++j) has[i] |= (lab[i * cols + j] + other_lab[i * cols + j]) != 0;
        any_unlabelled |= !has[i];
    }
    if (any_unlabelled) {
        std::vector<uint64_t> hp = P.down(), hd = D.down();
        for (size_t i = 0; i < rows; ++i)
            if (!has[i])
                for (size_t j = 0; j < cols; ++j) hd[i * cols + j] = hp[i * cols + j];
        D.up(hd);
    }
    detail::ideal_reshare(k, party, 0, &P, rows, cols, p);
    detail::ideal_reshare(k, party, 1, &D, rows, cols, p_minus_y);
}

// dstVec = in (.) ReLU'(z); g = dstVec * weightT unless this is the first layer (gcn.h:702-708)
inline void twoPartyGCNBackwardNNWithoutAH(const ShareVecVec& in, const ShareVecVec& z, const ShareTensor& weightT, ShareVecVec& dstVec,
                                           ShareVecVec& g, bool isFirstLayer, uint64_t coTid, int party) {
    using namespace cognn_shim;
    {
        Call k(coTid, party);
        size_t rows, cols, zr, zc;
        std::vector<uint64_t> x = flatten(in, &rows, &cols), zz = flatten(z, &zr, &zc), ox, oz;
        if (zr != rows || zc != cols) fatal("twoPartyGCNBackwardNNWithoutAH", "z shape differs from the gradient");
        if (!detail::ideal_exchange(k, party, x, ox)) {
            k.ch.send(zz);
            detail::ideal_reshare(k, party, 0, nullptr, rows, cols, dstVec);
        } else {
            k.ch.recv(oz);
            if (oz.size() != zz.size()) fatal("twoPartyGCNBackwardNNWithoutAH", "z size mismatch");
            Dev g0(k.c, x.size()), g1(k.c, x.size()), z0(k.c, x.size()), z1(k.c, x.size()), plain(k.c, x.size());
            g0.up(x);
            g1.up(ox);
            z0.up(zz);
            z1.up(oz);
            ok(k.c, cgb_ideal_relu_grad(k.c, g0.p, g1.p, z0.p, z1.p, plain.p, x.size()), "cgb_ideal_relu_grad");
            detail::ideal_reshare(k, party, 0, &plain, rows, cols, dstVec);
        }
    }
    if (!isFirstLayer) twoPartyGCNMatMul(dstVec, weightT, g, coTid, party);
    else g.clear();
}

// ---- the fused primitives of the UNOPTIMISED operators (original-gcn/gcn.h:459, 493, 586, 622; BASELINE configs[2]'s comparison
// arm): compositions of the primitives above, so they inherit the same 2PC-residual stand-ins.  `normalizer` is passed empty
// by the reference (original-gcn/gcn.h:451) and ignored.
// z = AH * W, new_h = ReLU(z)
inline void twoPartyGCNForwardNN(const ShareVecVec& ah, const ShareTensor& weight, const std::vector<uint64_t>& /*normalizer*/, ShareTensor& z,
                                 ShareTensor& new_h, uint64_t coTid, int party) {
    twoPartyGCNMatMul(ah, weight, z, coTid, party);
    twoPartyGCNRelu(z, new_h, coTid, party);
}
// z = AH * W, p = softmax(z), p_minus_y = p - label
inline void twoPartyGCNForwardNNPrediction(const ShareVecVec& ah, const ShareTensor& weight, const ShareVecVec& label,
                                           const std::vector<uint64_t>& /*normalizer*/, ShareTensor& z, ShareTensor& p, ShareTensor& p_minus_y,
                                           uint64_t coTid, int party) {
    twoPartyGCNMatMul(ah, weight, z, coTid, party);
    twoPartyGCNForwardNNPredictionWithoutWeight(z, label, p, p_minus_y, coTid, party);
}
// last layer, backward: d = AH^T * grad (weight gradient), g = grad * W^T (gradient towards the previous layer)
inline void twoPartyGCNBackwardNNInit(const ShareVecVec& grad, const ShareTensor& ah_t, const ShareTensor& weightT,
                                      const std::vector<uint64_t>& /*normalizer*/, ShareTensor& d, ShareTensor& g, uint64_t coTid, int party) {
    ShareTensor d_new, g_new;
    twoPartyGCNMatMul(ah_t, grad, d_new, coTid, party);
    twoPartyGCNMatMul(grad, weightT, g_new, coTid, party);
    d.swap(d_new);
    g.swap(g_new);
}
// hidden layer, backward: v = grad (.) ReLU'(z), d = AH^T * v, g = v * W^T unless this is the first layer
inline void twoPartyGCNBackwardNN(const ShareVecVec& grad, const ShareTensor& ah_t, const ShareTensor& z, const ShareTensor& weightT,
                                  const std::vector<uint64_t>& /*normalizer*/, ShareTensor& d, ShareTensor& g, bool isFirstLayer, uint64_t coTid,
                                  int party) {
    ShareVecVec v;
    ShareTensor g_new, d_new;
    twoPartyGCNBackwardNNWithoutAH(grad, z, weightT, v, g_new, isFirstLayer, coTid, party);
    twoPartyGCNMatMul(ah_t, v, d_new, coTid, party);
    d.swap(d_new);
    g.swap(g_new);
}
#endif  // COGNN_SHIM_IDEAL_NONLINEAR

// the per-edge scale of the UNOPTIMISED Scatter (original-gcn/gcn.h:243-250): out = in * n0[row] * n1[row], two fixed-point
// products; n0 (source out-degree normaliser) is ALICE's, n1 (destination in-degree normaliser) ALICE's for local edges and
// BOB's for mirror edges (ssk.h:1041-1043) -- each side passes zeros for what it does not know
inline void twoPartyGCNVectorScale(const ShareVecVec& in, const std::vector<uint64_t>& normalizer0, const std::vector<uint64_t>& normalizer1,
                                   ShareVecVec& out, uint64_t coTid, int party) {
    ShareVecVec tmp;
    detail::rowmul(in, normalizer0, tmp, coTid, party, SCALER_BIT_LENGTH);
    detail::rowmul(tmp, normalizer1, out, coTid, party, SCALER_BIT_LENGTH);
}

}  // namespace sci

// ---------------------------------------------------------------------------------------------------------------------
// OGA: group-by-destination sum over dst-sorted rows, result in the same "duplicated" layout (gcn.h:328-335).
// ALICE passes the real dstPos, BOB passes zeros (ssk.h:1047-1048) and never learns the groups: the map is linear, so
// BOB sends x1 - r, ALICE computes G(x0 + x1 - r) + (G r - s), BOB keeps s.
inline ShareVecVec prefix_network_aggregate(const std::vector<uint64_t>& dstPos, const ShareVecVec& svv, AggregationOp op, uint64_t coTid,
                                            int party, bool /*duplicate*/) {
    using namespace cognn_shim;
    if (op != AggregationOp::ADD_AGG) fatal("prefix_network_aggregate", "only ADD_AGG is on the GCN path");
    Call k(coTid, party);
    size_t E, D;
    std::vector<uint64_t> x = flatten(svv, &E, &D);
    cgb_ctx* c = k.c;
    ShareVecVec result;
    if (party == sci::BOB) {
        Dev dx(c, E * D), m(c, E * D), s(c, E * D);
        dx.up(x);
        ok(c, cgb_prg_mask_sub(c, k.rt.key, k.sid(K_OM_R), 0, dx.p, m.p, m.n), "cgb_prg_mask_sub");
        k.ch.send(m.down());
        prg(c, k.rt.key, k.sid(K_OM_S), s);
        unflatten(s.down(), E, D, result);
        return result;
    }
    if (dstPos.size() != E) fatal("prefix_network_aggregate", "one destination per row expected");
    std::vector<uint32_t> segptr(1, 0);
    for (size_t e = 1; e <= E; ++e)
        if (e == E || dstPos[e] != dstPos[e - 1]) segptr.push_back((uint32_t)e);
    if (E == 0) segptr.assign(1, 0);
    std::vector<uint64_t> msg;
    k.ch.recv(msg);
    if (msg.size() != E * D) fatal("prefix_network_aggregate", "message size mismatch");
    Dev dx(c, E * D), m(c, E * D), r(c, E * D), s(c, E * D), gr(c, E * D), y(c, E * D);
    void* dseg = nullptr;
    ok(c, cgb_malloc(c, segptr.size() * 4, &dseg), "cgb_malloc");
    ok(c, cgb_h2d(c, dseg, segptr.data(), segptr.size() * 4), "cgb_h2d");
    dx.up(x);
    m.up(msg);
    prg(c, k.rt.key, k.sid(K_OM_R), r);
    prg(c, k.rt.key, k.sid(K_OM_S), s);
    const uint32_t n_seg = (uint32_t)segptr.size() - 1;
    ok(c, cgb_segsum(c, (const uint32_t*)dseg, n_seg, E, r.p, gr.p, (uint32_t)D, 1), "cgb_segsum");  // dealer: G r
    ok(c, cgb_sub(c, gr.p, s.p, gr.p, gr.n), "cgb_sub");                                              //         - s
    ok(c, cgb_add(c, dx.p, m.p, dx.p, dx.n), "cgb_add");
    ok(c, cgb_segsum(c, (const uint32_t*)dseg, n_seg, E, dx.p, y.p, (uint32_t)D, 1), "cgb_segsum");
    ok(c, cgb_add(c, y.p, gr.p, y.p, y.n), "cgb_add");
    unflatten(y.down(), E, D, result);
    cgb_free(c, dseg);
    return result;
}

// OM online, client side (the party that knows the positions): dst[j] = src[index of dstPos[j] in srcPos] on shares
inline void client_oblivious_mapper_online(const std::vector<uint64_t>& srcPos, const std::vector<uint64_t>& dstPos, const ShareVecVec& srcSvv,
                                           ShareVecVec& dstSvv, uint32_t plainNumPerOperand, uint64_t iter, uint32_t preprocessId,
                                           uint64_t coTid, bool allowMissing = false) {
    using namespace cognn_shim;
    (void)iter;
    (void)preprocessId;  // the reference keys its offline correlation by (iter, preprocessId); here by the call counter
    Call k(coTid, sci::ALICE);
    const size_t n_src = srcPos.size(), n_dst = dstPos.size(), D = plainNumPerOperand;
    std::unordered_map<uint64_t, uint32_t> first;
    for (size_t i = 0; i < n_src; ++i) first.emplace(srcPos[i], (uint32_t)i);  // duplicated sources: first occurrence
    std::vector<uint32_t> idx(n_dst);
    for (size_t j = 0; j < n_dst; ++j) {
        auto it = first.find(dstPos[j]);
        if (it == first.end()) {
            if (!allowMissing) fatal("client_oblivious_mapper_online", "destination position missing in source positions");
            idx[j] = CGB_NO_ROW;
        } else idx[j] = it->second;
    }
    if (getenv("COGNN_SHIM_TRACE"))
        printf("[shim] OM client me=%d peer=%llu iter=%llu id=%u op=%llu n_src=%zu n_dst=%zu D=%zu\n", k.rt.tileIndex, (unsigned long long)coTid,
               (unsigned long long)iter, preprocessId, (unsigned long long)k.op, n_src, n_dst, D);
    k.ch.send(std::vector<uint64_t>{(uint64_t)n_dst});  // the server learns the output size (as in the reference's preprocessing)
    std::vector<uint64_t> msg;
    k.ch.recv(msg);
    // The client's own share is read only now, after the server's message: with more than two parties the ALICE threads of the
    // non-primary peers call this on gs.localVertexSvv while the primary thread may still be inside PreScatterComp
    // (ssk.h:734-763 has no barrier there); their servers answer only after the primary helper has forwarded its PreScatter result
    // (ssk.h:981-1002), which orders the two in practice.
    size_t r_, c_;
    std::vector<uint64_t> x = flatten(srcSvv, &r_, &c_);
    if (r_ != n_src || (n_src && c_ != D)) {
        char emsg[200];
        snprintf(emsg, sizeof emsg, "source shape mismatch: %zu positions, %zu x %zu shares, %zu columns expected (iter %llu, id %u, peer %llu)",
                 n_src, r_, c_, D, (unsigned long long)iter, preprocessId, (unsigned long long)coTid);
        fatal("client_oblivious_mapper_online", emsg);
    }
    if (msg.size() != n_src * D) fatal("client_oblivious_mapper_online", "message size mismatch");
    cgb_ctx* c = k.c;
    Dev dx(c, n_src * D), m(c, n_src * D), r(c, n_src * D), s(c, n_dst * D), delta(c, n_dst * D), y(c, n_dst * D);
    void* didx = nullptr;
    ok(c, cgb_malloc(c, (n_dst ? n_dst : 1) * 4, &didx), "cgb_malloc");
    if (n_dst) ok(c, cgb_h2d(c, didx, idx.data(), n_dst * 4), "cgb_h2d");
    dx.up(x);
    m.up(msg);
    prg(c, k.rt.key, k.sid(K_OM_R), r);
    prg(c, k.rt.key, k.sid(K_OM_S), s);
    if (D) {
        ok(c, cgb_expand_rows(c, (const uint32_t*)didx, n_dst, r.p, nullptr, delta.p, (uint32_t)D), "cgb_expand_rows");  // dealer: pi(r)
        ok(c, cgb_sub(c, delta.p, s.p, delta.p, delta.n), "cgb_sub");                                                   //         - s
        ok(c, cgb_add(c, dx.p, m.p, dx.p, dx.n), "cgb_add");
        ok(c, cgb_expand_rows(c, (const uint32_t*)didx, n_dst, dx.p, delta.p, y.p, (uint32_t)D), "cgb_expand_rows");
    }
    unflatten(y.down(), n_dst, D, dstSvv);
    cgb_free(c, didx);
}

// OM online, server side: sends its masked share, keeps the fresh mask as its share of the result
inline void server_oblivious_mapper_online(const ShareVecVec& srcSvv, ShareVecVec& dstSvv, uint64_t iter, uint32_t preprocessId, uint64_t coTid) {
    using namespace cognn_shim;
    (void)iter;
    (void)preprocessId;
    Call k(coTid, sci::BOB);
    size_t n_src, D;
    std::vector<uint64_t> x = flatten(srcSvv, &n_src, &D);
    std::vector<uint64_t> hdr;
    k.ch.recv(hdr);
    const size_t n_dst = hdr.empty() ? 0 : (size_t)hdr[0];
    if (getenv("COGNN_SHIM_TRACE"))
        printf("[shim] OM server me=%d owner=%llu iter=%llu id=%u op=%llu n_src=%zu n_dst=%zu D=%zu\n", k.rt.tileIndex, (unsigned long long)coTid,
               (unsigned long long)iter, preprocessId, (unsigned long long)k.op, n_src, n_dst, D);
    cgb_ctx* c = k.c;
    Dev dx(c, n_src * D), m(c, n_src * D), s(c, n_dst * D);
    dx.up(x);
    ok(c, cgb_prg_mask_sub(c, k.rt.key, k.sid(K_OM_R), 0, dx.p, m.p, m.n), "cgb_prg_mask_sub");
    k.ch.send(m.down());
    prg(c, k.rt.key, k.sid(K_OM_S), s);
    unflatten(s.down(), n_dst, D, dstSvv);
}
