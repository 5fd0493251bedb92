// comm.cpp -- message planes for the engine: LoopbackComm (all parties in one process, device-to-device copies) and
// NcclComm (one party per process / GPU: ncclSend / ncclRecv grouped per protocol round over NVLink).  They replace the
// reference's host TCP planes (CommSync over osuCrypto::Channel, include/comm_sync.h:212-277, engine.h:157-201; TaskComm).
// Message bytes are the raw little-endian row-major u64 buffer (no Boost archive framing).
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstring>
#include <deque>
#include <map>
#include <stdexcept>

#include "engine.h"

namespace cognn {

namespace {

struct Post {
    int src, dst;
    uint64_t* p;
    size_t n;
    std::string tag;
};

class LoopbackComm : public Comm {
public:
    LoopbackComm(int world, cgb_ctx* ctx) : world_(world), ctx_(ctx) {}
    int world() const override { return world_; }
    bool is_local(int) const override { return true; }
    cgb_ctx* ctx() override { return ctx_; }
    void post_send(int src, int dst, const uint64_t* d, size_t n, const std::string& tag) override {
        sends_[{src, dst}].push_back(Post{src, dst, const_cast<uint64_t*>(d), n, tag});
        order_.push_back({src, dst});
    }
    void post_recv(int dst, int src, uint64_t* d, size_t n) override {
        recvs_[{src, dst}].push_back(Post{src, dst, d, n, ""});
    }
    void exchange() override {
        ++rounds;
        std::vector<uint64_t*> dst;
        std::vector<const uint64_t*> src;
        std::vector<uint64_t> len;
        for (auto& pr : order_) {
            auto& sq = sends_[pr];
            auto& rq = recvs_[pr];
            if (sq.empty()) continue;
            if (rq.empty()) throw std::runtime_error("LoopbackComm: send without matching recv (" + sq.front().tag + ")");
            Post s = sq.front(), r = rq.front();
            sq.pop_front();
            rq.pop_front();
            if (s.n != r.n) throw std::runtime_error("LoopbackComm: size mismatch on " + s.tag);
            dst.push_back(r.p);
            src.push_back(s.p);
            len.push_back(s.n);
            words_sent += s.n;
            if (record) {
                Message m{cur_iter, s.src, s.dst, s.tag, std::vector<uint64_t>(s.n)};
                if (s.n) cgb_d2h(ctx_, m.data.data(), s.p, s.n * sizeof(uint64_t));
                cgb_ctx_sync(ctx_);
                transcript.push_back(std::move(m));
            }
        }
        // every message of the round in one launch (one copy node per message would serialise on the stream)
        if (!dst.empty() && cgb_copy_segments(ctx_, dst.data(), src.data(), len.data(), (uint32_t)dst.size()) != CGB_OK)
            throw std::runtime_error(std::string("LoopbackComm: copy failed: ") + cgb_last_error(ctx_));
        order_.clear();
        for (auto& kv : recvs_)
            if (!kv.second.empty()) throw std::runtime_error("LoopbackComm: recv without matching send");
    }

private:
    int world_;
    cgb_ctx* ctx_;
    std::map<std::pair<int, int>, std::deque<Post>> sends_, recvs_;
    std::vector<std::pair<int, int>> order_;
};

#define NCCL_CK(call)                                                                              \
    do {                                                                                           \
        ncclResult_t _r = (call);                                                                  \
        if (_r != ncclSuccess) throw std::runtime_error(std::string(#call) + ": " + ncclGetErrorString(_r)); \
    } while (0)

class NcclComm : public Comm {
public:
    NcclComm(int rank, int world, cgb_ctx* ctx, const void* uid) : rank_(rank), world_(world), ctx_(ctx) {
        ncclUniqueId id;
        memcpy(&id, uid, sizeof(id));
        NCCL_CK(ncclCommInitRank(&comm_, world, id, rank));
    }
    ~NcclComm() override {
        if (d_flag_) cudaFree(d_flag_);
        if (comm_) ncclCommDestroy(comm_);
    }
    void barrier() override {
        cudaStream_t st = (cudaStream_t)cgb_ctx_stream(ctx_);
        if (!d_flag_) {
            if (cudaMalloc(&d_flag_, sizeof(int)) != cudaSuccess) throw std::runtime_error("NcclComm: cudaMalloc failed");
            cudaMemsetAsync(d_flag_, 0, sizeof(int), st);
        }
        NCCL_CK(ncclAllReduce(d_flag_, d_flag_, 1, ncclInt, ncclSum, comm_, st));
        if (cudaStreamSynchronize(st) != cudaSuccess) throw std::runtime_error("NcclComm: barrier failed");
    }
    int world() const override { return world_; }
    bool is_local(int p) const override { return p == rank_; }
    cgb_ctx* ctx() override { return ctx_; }
    void post_send(int src, int dst, const uint64_t* d, size_t n, const std::string& tag) override {
        if (src != rank_) throw std::runtime_error("NcclComm: send from a party that is not hosted here");
        posts_.push_back(Post{src, dst, const_cast<uint64_t*>(d), n, tag});
    }
    void post_recv(int dst, int src, uint64_t* d, size_t n) override {
        if (dst != rank_) throw std::runtime_error("NcclComm: recv for a party that is not hosted here");
        posts_.push_back(Post{src, dst, d, n, ""});
    }
    void exchange() override {
        ++rounds;
        cudaStream_t st = (cudaStream_t)cgb_ctx_stream(ctx_);
        // self-sends (T == 1) are plain copies, matched in order
        std::deque<Post> self_s, self_r;
        NCCL_CK(ncclGroupStart());
        for (auto& p : posts_) {
            if (p.src == p.dst) {
                (p.tag.empty() ? self_r : self_s).push_back(p);
                continue;
            }
            if (p.src == rank_) {
                if (p.n) NCCL_CK(ncclSend(p.p, p.n, ncclUint64, p.dst, comm_, st));
                words_sent += p.n;
            } else {
                if (p.n) NCCL_CK(ncclRecv(p.p, p.n, ncclUint64, p.src, comm_, st));
            }
        }
        NCCL_CK(ncclGroupEnd());
        while (!self_s.empty() && !self_r.empty()) {
            cgb_d2d(ctx_, self_r.front().p, self_s.front().p, self_s.front().n * sizeof(uint64_t));
            self_s.pop_front();
            self_r.pop_front();
        }
        if (record) {
            for (auto& p : posts_) {
                if (p.src != rank_ || p.tag.empty()) continue;
                Message m{cur_iter, p.src, p.dst, p.tag, std::vector<uint64_t>(p.n)};
                if (p.n) cgb_d2h(ctx_, m.data.data(), p.p, p.n * sizeof(uint64_t));
                cgb_ctx_sync(ctx_);
                transcript.push_back(std::move(m));
            }
        }
        posts_.clear();
    }

private:
    int rank_, world_;
    cgb_ctx* ctx_;
    ncclComm_t comm_ = nullptr;
    int* d_flag_ = nullptr;
    std::vector<Post> posts_;
};

}  // namespace

std::unique_ptr<Comm> make_loopback_comm(int world, cgb_ctx* ctx) { return std::unique_ptr<Comm>(new LoopbackComm(world, ctx)); }
std::unique_ptr<Comm> make_nccl_comm(int rank, int world, cgb_ctx* ctx, const void* uid) {
    return std::unique_ptr<Comm>(new NcclComm(rank, world, ctx, uid));
}
void nccl_get_unique_id(void* out128) {
    ncclUniqueId id;
    NCCL_CK(ncclGetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
}

}  // namespace cognn
