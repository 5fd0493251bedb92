// cognn_taskworker.h -- what the reference's engine and harness expect from the absent Task-Worker tree BESIDES the arithmetic
// primitives of cognn_shim.h: TaskComm (per-peer host channels, role semaphores, run flags), Semaphore, TaskqHandlerConfig,
// print_duration, get_next_power_of_2, the oblivious-mapper preprocessing entry points and the SCI channel set-up.  With these
// (and stand-ins for Boost / cryptoTools) /root/reference/algo_kernels/common_harness/harness.cpp, include/*.h and the GCN
// operator headers compile UNCHANGED and run on the B200 library: tests/test_gpu_reference_dropin.py.
//
// Names, argument order and behaviour are inferred from the call sites (cited per item); one party per process, like the
// reference (`-t T -i me`).  Host transport: cognn_shim_net.h (TCP, loopback unless COGNN_SHIM_HOST_<i> names party i's host).
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "../cognn_shim.h"
#include "../cognn_shim_net.h"

// ---- small utilities ---------------------------------------------------------------------------------------------------
// ssk.h:192, 244, 612, 745, ...: "::<tag> took <seconds>" lines that tools/plot/*.py parse
template <typename TimePoint>
inline void print_duration(const TimePoint& t0, const char* tag) {
    const double s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    printf("::%s took %lf seconds\n", tag, s);
    fflush(stdout);
}
template <typename TimePoint>
inline void print_duration(const TimePoint& t0, const std::string& tag) { print_duration(t0, tag.c_str()); }
// ssk.h:369, 391 (power-of-two padding of a destination's source list without -r)
inline uint64_t get_next_power_of_2(uint64_t n) {
    uint64_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

// ssk.h:838-841, 1069-1072; gcn.h:747-792
class Semaphore {
public:
    explicit Semaphore(int count = 0) : count_(count) {}
    void release() {
        {
            std::lock_guard<std::mutex> l(m_);
            ++count_;
        }
        cv_.notify_one();
    }
    void acquire() {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return count_ > 0; });
        --count_;
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    int count_;
};

// ssk.h:782-785, 931-934, 1029: knobs of the HE task queue; kept as plain data, nothing on the GCN path reads them
struct TaskqHandlerConfig {
    bool sendOperand = false, mergeResult = false, sendTaskqDigest = false;
    int rotation = 0;
};

// include/algo_kernel.h:103-105 and include/vertex_centric_algo_kernel.h:103-105: entry points of the HE task queue of the
// older (non secret-shared) kernels.  The secret-shared GCN kernel overrides runAlgoKernelServer (ssk.h:912) and never queues a
// task, so these only have to exist.
inline void task_queue_handler(std::queue<Task>*) {
    cognn_shim::fatal("task_queue_handler", "the HE task queue is not part of the secret-shared GCN path (out of scope)");
}
inline void issue_server_recv_threads(std::vector<std::thread>&) {}

namespace cognn_shim {

// host channel of the shim's two-party primitives over a socket connection
struct NetChannel : Channel {
    std::unique_ptr<net::Conn> conn;
    explicit NetChannel(net::Conn* c) : conn(c) {}
    void send(const std::vector<uint64_t>& v) override { conn->send(std::string((const char*)v.data(), v.size() * 8)); }
    void recv(std::vector<uint64_t>& v) override {
        std::string s;
        conn->recv(s);
        v.resize(s.size() / 8);
        if (!s.empty()) memcpy(v.data(), s.data(), v.size() * 8);
    }
};
inline std::string host_of(size_t party) {
    const std::string key = "COGNN_SHIM_HOST_" + std::to_string(party);
    const char* e = getenv(key.c_str());
    return e ? e : "127.0.0.1";
}
inline void put_svv(net::Conn& c, const ShareVecVec& m) {
    std::string s;
    const uint64_t rows = m.size();
    s.append((const char*)&rows, 8);
    for (const auto& r : m) {
        const uint64_t n = r.size();
        s.append((const char*)&n, 8);
        s.append((const char*)r.data(), n * 8);
    }
    c.send(std::move(s));
}
inline void get_svv(net::Conn& c, ShareVecVec& m) {
    std::string s;
    c.recv(s);
    size_t o = 0;
    uint64_t rows = 0;
    memcpy(&rows, s.data(), 8);
    o += 8;
    ShareVecVec out(rows);
    for (auto& r : out) {
        uint64_t n = 0;
        memcpy(&n, s.data() + o, 8);
        o += 8;
        r.resize(n);
        if (n) memcpy(r.data(), s.data() + o, n * 8);
        o += n * 8;
    }
    m.swap(out);
}

}  // namespace cognn_shim

// ---- TaskComm -------------------------------------------------------------------------------------------------------------
// Two instances per process (harness.cpp:126-153): the CLIENT instance of party a owns one duplex channel to the SERVER
// instance of every other party b.  ssk.h:209-232 sends shares client -> server; the weight averaging of gcn.h:753-777 also
// answers server -> client on the same channel.
class TaskComm {
public:
    static TaskComm& getClientInstance() {
        static TaskComm c(true);
        return c;
    }
    static TaskComm& getServerInstance() {
        static TaskComm s(false);
        return s;
    }
    void tileNumIs(size_t n) { tileNum_ = n; }
    void tileIndexIs(size_t i) { tileIndex_ = i; }
    void settingIs(const std::string& s) { setting_ = s; }
    void noPreprocessIs(bool b) { noPreprocess_ = b; }
    void isClusterIs(bool b) { isCluster_ = b; }
    void isNoDummyEdgeIs(bool b) { isNoDummyEdge_ = b; }
    size_t getTileNum() const { return tileNum_; }
    size_t getTileIndex() const { return tileIndex_; }
    const std::string& getSetting() const { return setting_; }
    bool getNoPreprocess() const { return noPreprocess_; }
    bool getIsCluster() const { return isCluster_; }
    bool getIsNoDummyEdge() const { return isNoDummyEdge_; }

    // harness.cpp:144-153: both instances are set up concurrently in every process
    void setUp(bool isClient) {
        const int base = cognn_shim::net::port_base();
        conns_.resize(tileNum_);
        taskv_.resize(tileNum_);
        thc_.resize(tileNum_);
        localUpdateReady_.clear();
        remoteUpdateReady_.clear();
        for (size_t i = 0; i < tileNum_; ++i) {
            localUpdateReady_.emplace_back(new Semaphore(0));
            remoteUpdateReady_.emplace_back(new Semaphore(0));
        }
        for (size_t i = 0; i < tileNum_; ++i) {
            if (i == tileIndex_) continue;
            if (isClient)  // my client -> party i's server
                conns_[i].reset(cognn_shim::net::connect_named(cognn_shim::host_of(i), base + (int)i,
                                                               "tc:" + std::to_string(tileIndex_) + "->" + std::to_string(i)));
            else           // party i's client -> my server
                conns_[i].reset(cognn_shim::net::accept_named(base + (int)tileIndex_,
                                                              "tc:" + std::to_string(i) + "->" + std::to_string(tileIndex_)));
        }
    }
    void closeChannels() {
        for (auto& c : conns_)
            if (c) c->close();
    }
    void sendShareVecVec(const ShareVecVec& m, size_t peer) { cognn_shim::put_svv(*conns_.at(peer), m); }
    void recvShareVecVec(ShareVecVec& m, size_t peer) { cognn_shim::get_svv(*conns_.at(peer), m); }
    void sendShareTensorVec(const ShareTensorVec& tv, size_t peer) {  // ssk.h:231
        ShareVecVec flat;
        ShareVec hdr(1, tv.size());
        for (const auto& t : tv) hdr.push_back(t.size());
        flat.push_back(hdr);
        for (const auto& t : tv) flat.insert(flat.end(), t.begin(), t.end());
        sendShareVecVec(flat, peer);
    }
    void recvShareTensorVec(ShareTensorVec& tv, size_t peer) {  // ssk.h:232
        ShareVecVec flat;
        recvShareVecVec(flat, peer);
        ShareTensorVec out(flat.at(0).at(0));
        size_t row = 1;
        for (size_t k = 0; k < out.size(); ++k) {
            const size_t n = flat[0][1 + k];
            out[k].assign(flat.begin() + row, flat.begin() + row + n);
            row += n;
        }
        tv.swap(out);
    }
    void sendFinish() {  // ssk.h:270: every server tells its clients that it is done
        for (size_t i = 0; i < conns_.size(); ++i)
            if (conns_[i]) conns_[i]->send(std::string("FIN"));
    }
    void recvFinish() {  // ssk.h:272
        for (size_t i = 0; i < conns_.size(); ++i)
            if (conns_[i]) {
                std::string s;
                conns_[i]->recv(s);
            }
    }
    TaskqHandlerConfig& getTaskqHandlerConfig(size_t i) { return thc_.at(i); }
    std::vector<Task>& getTaskv(size_t i) { return taskv_.at(i); }
    Semaphore& getLocalUpdateReadySmp(size_t i) { return *localUpdateReady_.at(i); }
    Semaphore& getRemoteUpdateReadySmp(size_t i) { return *remoteUpdateReady_.at(i); }
    Semaphore& getRemoteWeightReadySmp() { return remoteWeightReady_; }
    Semaphore& getWeightAvgFinishedSmp() { return weightAvgFinished_; }

private:
    explicit TaskComm(bool isClient) : isClient_(isClient) {}
    bool isClient_;
    size_t tileNum_ = 1, tileIndex_ = 0;
    std::string setting_;
    bool noPreprocess_ = false, isCluster_ = false, isNoDummyEdge_ = false;
    std::vector<std::unique_ptr<cognn_shim::net::Conn>> conns_;
    std::vector<std::vector<Task>> taskv_;
    std::vector<TaskqHandlerConfig> thc_;
    std::vector<std::unique_ptr<Semaphore>> localUpdateReady_, remoteUpdateReady_;
    Semaphore remoteWeightReady_{0}, weightAvgFinished_{0};
};

// ---- SCI channel set-up (harness.cpp:155-168) ------------------------------------------------------------------------------
// setUpSCIChannel creates the party's runtime (one per process), set_up_mpc_channel(isClient, i) the host channel of the
// (me as ALICE, i as BOB) pair or of the (i as ALICE, me as BOB) pair.  The master key of the dealer emulation comes from
// COGNN_SHIM_KEY (eight 32-bit words, hex, comma separated) and must be the same for all parties of a run.
namespace sci {
inline void setUpSCIChannel() {
    TaskComm& tc = TaskComm::getClientInstance();
    uint32_t key[8] = {45, 0, 0, 0, 0, 0, 0, 0};
    if (const char* e = getenv("COGNN_SHIM_KEY")) {
        unsigned long long v[8] = {0};
        if (sscanf(e, "%llx,%llx,%llx,%llx,%llx,%llx,%llx,%llx", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]) >= 1)
            for (int i = 0; i < 8; ++i) key[i] = (uint32_t)v[i];
    }
    const char* dev = getenv("COGNN_B200_DEVICE");
    static cognn_shim::Runtime* rt = nullptr;
    if (!rt) {
        printf("cognn_shim: INSECURE test configuration -- trusted-dealer emulation for all correlated randomness%s\n",
#ifdef COGNN_SHIM_IDEAL_NONLINEAR
               " and ideal-functionality stand-ins for ReLU / softmax / ReLU' (COGNN_SHIM_IDEAL_NONLINEAR)"
#else
               ""
#endif
        );
        rt = new cognn_shim::Runtime((int)tc.getTileIndex(), (int)tc.getTileNum(), dev ? atoi(dev) : 0, key);
        cognn_shim::Runtime::bind_process(rt);
    }
}
}  // namespace sci
inline void set_up_mpc_channel(bool isClient, uint32_t i) {
    cognn_shim::Runtime& rt = cognn_shim::Runtime::current();
    const int base = cognn_shim::net::port_base();
    const std::string me = std::to_string(rt.tileIndex), other = std::to_string(i);
    cognn_shim::net::Conn* c = isClient ? cognn_shim::net::connect_named(cognn_shim::host_of(i), base + (int)i, "mpc:" + me + "->" + other)
                                        : cognn_shim::net::accept_named(base + rt.tileIndex, "mpc:" + other + "->" + me);
    rt.connect_owned(i, isClient ? 1 : 2, new cognn_shim::NetChannel(c));
}

// ---- oblivious mapper preprocessing (ssk.h:559-600 client, 637-670 server) ---------------------------------------------------
// The reference generates, per (iter, preprocessId, peer), the permutation correlation the online phase consumes.  Here the
// correlation comes from the trusted-dealer emulation inside the online call (cognn_shim.h), so preprocessing has nothing to
// exchange; it reports the number of iterations one batch covers (the whole epoch).
inline uint64_t client_gcn_batch_oblivious_mapper_preprocess(const std::vector<uint64_t>& /*srcPos*/, const std::vector<uint64_t>& /*dstPos*/,
                                                             const std::vector<uint32_t>& dimensions, uint32_t /*iter*/,
                                                             uint32_t /*preprocessId*/, uint64_t /*coTid*/, bool /*allowMissing*/ = false) {
    return dimensions.empty() ? 1 : dimensions.size();
}
inline uint64_t server_gcn_batch_oblivious_mapper_preprocess(const std::vector<uint32_t>& dimensions, uint32_t /*iter*/,
                                                             uint32_t /*preprocessId*/, uint64_t /*coTid*/) {
    return dimensions.empty() ? 1 : dimensions.size();
}
