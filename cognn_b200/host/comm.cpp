// comm.cpp -- message planes for the engine: LoopbackComm (all parties in one process, one segmented copy launch per round) and
// NcclComm (one party per process / GPU: ncclSend / ncclRecv grouped per protocol round over NVLink, or -- after reserve() --
// peer-memory rounds: pushes into the receiver's slot and flags, cgb_peer_round).  They replace the
// reference's host TCP planes (CommSync over osuCrypto::Channel, include/comm_sync.h:212-277, engine.h:157-201; TaskComm).
// Message bytes are the raw little-endian row-major u64 buffer (no Boost archive framing).
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
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
        for (auto& l : links_)
            if (l.peer_base) cgb_ipc_close(ctx_, l.peer_base);
        if (block_) cudaFree(block_);
        if (state_) cudaFree(state_);
        if (d_flag_) cudaFree(d_flag_);
        if (comm_) ncclCommDestroy(comm_);
    }
    const char* plane() const override { return peer_ok_ ? "peer-memory rounds over NVLink (nccl bootstrap / fallback)" : "nccl"; }

    // ---- peer-memory plane --------------------------------------------------------------------------------------------
    // Every ordered pair of parties gets a double-buffered slot pair in the receiver's memory, mapped by the sender with CUDA
    // IPC, plus a data flag (receiver side) and an acknowledge flag (sender side).  A round is then ONE launch per rank
    // (cgb_peer_round: the pushes first, then the consumption of what arrives) instead of an NCCL group: no proxy thread, no
    // channel set-up, and the round counters live on the device so the launches replay from a CUDA graph.  Handles travel once, over
    // NCCL.  Pairs whose messages do not fit the slot (or more than 16 messages) keep using ncclSend / ncclRecv, and so does
    // everything when any rank could not map its peers (different nodes, IPC disabled) or when the plane is off (see below).
    void reserve(size_t max_words) override {
        if (world_ < 2 || peer_ok_ || max_words == 0) return;
        // Default: on for more than 4 parties.  Measured: 8 parties 4.5 ms against 6.0 ms per arxiv-shaped
        // epoch over ncclSend / ncclRecv (rounds with six peers), a tie at 2, and 8 % behind NCCL at 4 parties on the
        // CiteSeer shape, whose 25 MB product messages pay for the copy out of the slot.  =1 / =0 force it on / off.
        const char* env = getenv("COGNN_B200_PEER_EXCHANGE");
        int ok = (env ? env[0] != '0' : world_ > 4) && world_ <= 16;
        cudaStream_t st = (cudaStream_t)cgb_ctx_stream(ctx_);
        const size_t slot = (max_words + 1) & ~(size_t)1;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if ((size_t)world_ * 2 * slot * 8 > free_b / 4) ok = 0;  // arenas would take more than a quarter of what is free
        links_.assign(world_, Link{});
        // ONE allocation per rank -- [4 KB flag page | two slots for every source party] -- so that a peer maps one handle
        const size_t blob = 64;
        const size_t block_words = 512 + (size_t)world_ * 2 * slot;
        std::vector<unsigned char> mine(blob, 0), all(blob * world_, 0);
        if (ok) {
            ok = cudaMalloc((void**)&block_, block_words * 8) == cudaSuccess &&
                 cudaMalloc((void**)&state_, (4 * world_ + 1) * sizeof(uint32_t)) == cudaSuccess;
            if (ok) {
                cudaMemsetAsync(block_, 0, 4096, st);
                cudaMemsetAsync(state_, 0, (4 * world_ + 1) * sizeof(uint32_t), st);
                flags_ = (uint32_t*)block_;
                ok = cgb_ipc_export(ctx_, block_, mine.data()) == CGB_OK;
            }
            for (int p = 0; ok && p < world_; ++p) links_[p].arena_in = block_ + 512 + (size_t)p * 2 * slot;
        }
        if (!all_agree(ok)) return release_peer();
        // handles of everybody, through NCCL (device staging)
        unsigned char *d_mine = nullptr, *d_all = nullptr;
        if (cudaMalloc((void**)&d_mine, blob) != cudaSuccess || cudaMalloc((void**)&d_all, blob * world_) != cudaSuccess)
            throw std::runtime_error("NcclComm: cudaMalloc failed");
        cudaMemcpyAsync(d_mine, mine.data(), blob, cudaMemcpyHostToDevice, st);
        NCCL_CK(ncclAllGather(d_mine, d_all, blob, ncclUint8, comm_, st));
        cudaMemcpyAsync(all.data(), d_all, blob * world_, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        cudaFree(d_mine);
        cudaFree(d_all);
        for (int p = 0; ok && p < world_; ++p) {
            if (p == rank_) continue;
            void* base = nullptr;
            ok = cgb_ipc_open(ctx_, all.data() + blob * (size_t)p, &base) == CGB_OK;
            links_[p].peer_base = (uint64_t*)base;
            if (ok) {
                links_[p].peer_flags = (uint32_t*)base;
                links_[p].arena_out = (uint64_t*)base + 512 + (size_t)rank_ * 2 * slot;
            }
        }
        if (!all_agree(ok)) return release_peer();
        slot_words_ = slot;
        peer_ok_ = true;
    }

    void check_peer_errors() {
        if (!peer_ok_) return;
        uint32_t e = 0;
        cudaMemcpy(&e, state_ + 4 * world_, sizeof(e), cudaMemcpyDeviceToHost);
        if (e) throw std::runtime_error(e == 2 ? "NcclComm: peer round timed out waiting for an acknowledgement"
                                               : "NcclComm: peer round timed out waiting for data");
    }
    void barrier() override {
        cudaStream_t st = (cudaStream_t)cgb_ctx_stream(ctx_);
        if (!d_flag_) {
            if (cudaMalloc(&d_flag_, sizeof(int)) != cudaSuccess) throw std::runtime_error("NcclComm: cudaMalloc failed");
            cudaMemsetAsync(d_flag_, 0, sizeof(int), st);
        }
        NCCL_CK(ncclAllReduce(d_flag_, d_flag_, 1, ncclInt, ncclSum, comm_, st));
        if (cudaStreamSynchronize(st) != cudaSuccess) throw std::runtime_error("NcclComm: barrier failed");
        check_peer_errors();
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
        // which pairs run over peer memory this round: both ends see the same message count and sizes, so they decide alike
        std::vector<size_t> out_words(world_, 0), in_words(world_, 0);
        std::vector<int> out_cnt(world_, 0), in_cnt(world_, 0);
        auto even = [](size_t n) { return (n + 1) & ~(size_t)1; };
        for (auto& p : posts_) {
            if (p.src == p.dst || p.n == 0) continue;
            if (p.src == rank_) { out_words[p.dst] += even(p.n); ++out_cnt[p.dst]; }
            else { in_words[p.src] += even(p.n); ++in_cnt[p.src]; }
        }
        auto via_peer = [&](size_t words, int cnt) { return peer_ok_ && cnt > 0 && cnt <= 16 && words <= slot_words_; };
        if (peer_ok_) peer_launches(out_words, out_cnt, in_words, in_cnt, via_peer);
        // self-sends (T == 1) are plain copies, matched in order; everything not taken by the peer plane goes through NCCL
        std::deque<Post> self_s, self_r;
        bool any_nccl = false;
        for (auto& p : posts_) {
            if (p.src == p.dst) continue;
            const bool out = p.src == rank_;
            if (p.n && !via_peer(out ? out_words[p.dst] : in_words[p.src], out ? out_cnt[p.dst] : in_cnt[p.src])) any_nccl = true;
        }
        if (any_nccl) NCCL_CK(ncclGroupStart());
        for (auto& p : posts_) {
            if (p.src == p.dst) {
                (p.tag.empty() ? self_r : self_s).push_back(p);
                continue;
            }
            if (p.src == rank_) {
                if (p.n && !via_peer(out_words[p.dst], out_cnt[p.dst])) NCCL_CK(ncclSend(p.p, p.n, ncclUint64, p.dst, comm_, st));
                words_sent += p.n;
            } else {
                if (p.n && !via_peer(in_words[p.src], in_cnt[p.src])) NCCL_CK(ncclRecv(p.p, p.n, ncclUint64, p.src, comm_, st));
            }
        }
        if (any_nccl) NCCL_CK(ncclGroupEnd());
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
    struct Link {
        uint64_t* arena_in = nullptr;    // local: two slots the peer pushes into
        uint64_t* peer_base = nullptr;   // the peer's block, mapped with CUDA IPC
        uint64_t* arena_out = nullptr;   // inside it: my two slots
        uint32_t* peer_flags = nullptr;  // inside it: the peer's flag page
    };
    // flag page (local, written by peers): [p] = data flag of source p, [world + p] = acknowledge flag of destination p
    // state (local): [p] send seq, [world + p] recv seq, [2 world + p] send done, [3 world + p] recv done, [4 world] error word
    bool all_agree(int ok) {
        cudaStream_t st = (cudaStream_t)cgb_ctx_stream(ctx_);
        int* d = nullptr;
        if (cudaMalloc((void**)&d, sizeof(int)) != cudaSuccess) throw std::runtime_error("NcclComm: cudaMalloc failed");
        cudaMemcpyAsync(d, &ok, sizeof(int), cudaMemcpyHostToDevice, st);
        NCCL_CK(ncclAllReduce(d, d, 1, ncclInt, ncclMin, comm_, st));
        int r = 0;
        cudaMemcpyAsync(&r, d, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        cudaFree(d);
        return r != 0;
    }
    void release_peer() {
        cudaGetLastError();
        for (auto& l : links_) {
            if (l.peer_base) cgb_ipc_close(ctx_, l.peer_base);
            l = Link{};
        }
        if (block_) cudaFree(block_);
        if (state_) cudaFree(state_);
        block_ = nullptr;
        flags_ = state_ = nullptr;
        peer_ok_ = false;
    }
    // One launch per round when it fits (16 segments, 16 links): the pushes of every outgoing pair that runs over peer memory,
    // then the consumption of every incoming one.  Otherwise whole links are spread over several launches, pushes first.
    template <typename Fits>
    void peer_launches(const std::vector<size_t>& out_words, const std::vector<int>& out_cnt, const std::vector<size_t>& in_words,
                       const std::vector<int>& in_cnt, Fits&& via_peer) {
        std::vector<cgb_xseg> segs;
        std::vector<cgb_xlink> lks;
        size_t biggest = 0;
        auto flush = [&]() {
            if (segs.empty()) return;
            const uint32_t ctas = (uint32_t)std::min<size_t>(64, std::max<size_t>(1, (biggest + 8191) / 8192));
            if (cgb_peer_round(ctx_, segs.data(), (uint32_t)segs.size(), lks.data(), (uint32_t)lks.size(), slot_words_, ctas,
                               state_ + 4 * world_) != CGB_OK)
                throw std::runtime_error(std::string("NcclComm: cgb_peer_round failed: ") + cgb_last_error(ctx_));
            segs.clear();
            lks.clear();
            biggest = 0;
        };
        for (int mode = 0; mode < 2; ++mode) {
            const std::vector<size_t>& words = mode == 0 ? out_words : in_words;
            const std::vector<int>& cnt = mode == 0 ? out_cnt : in_cnt;
            for (int peer = 0; peer < world_; ++peer) {
                if (peer == rank_ || !via_peer(words[peer], cnt[peer])) continue;
                if (segs.size() + (size_t)cnt[peer] > 16 || lks.size() == 16) flush();
                cgb_xlink l;
                l.recv = (uint32_t)mode;
                if (mode == 0) {
                    l.seq = state_ + peer;
                    l.done = state_ + 2 * world_ + peer;
                    l.wait_flag = flags_ + world_ + peer;                      // the peer's acknowledgements of what I sent it
                    l.signal_flag = links_[peer].peer_flags + rank_;           // its data flag for source = me
                } else {
                    l.seq = state_ + world_ + peer;
                    l.done = state_ + 3 * world_ + peer;
                    l.wait_flag = flags_ + peer;                               // data flag of source = peer
                    l.signal_flag = links_[peer].peer_flags + world_ + rank_;  // its acknowledge flag for destination = me
                }
                size_t off = 0;
                for (auto& p : posts_) {
                    if (p.src == p.dst || p.n == 0) continue;
                    if (mode == 0 ? (p.src != rank_ || p.dst != peer) : (p.dst != rank_ || p.src != peer)) continue;
                    cgb_xseg sg;
                    if (mode == 0) {
                        sg.src = p.p;
                        sg.dst = links_[peer].arena_out + off;
                    } else {
                        sg.src = links_[peer].arena_in + off;
                        sg.dst = p.p;
                    }
                    sg.n_words = p.n;
                    sg.link = (uint32_t)lks.size();
                    segs.push_back(sg);
                    biggest = std::max(biggest, p.n);
                    off += (p.n + 1) & ~(size_t)1;
                }
                lks.push_back(l);
            }
        }
        flush();
    }
    int rank_, world_;
    cgb_ctx* ctx_;
    ncclComm_t comm_ = nullptr;
    int* d_flag_ = nullptr;
    std::vector<Post> posts_;
    std::vector<Link> links_;
    uint64_t* block_ = nullptr;   // [4 KB flag page | 2 slots per source party], exported to every peer
    uint32_t* flags_ = nullptr;   // = block_
    uint32_t* state_ = nullptr;
    size_t slot_words_ = 0;
    bool peer_ok_ = false;
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
