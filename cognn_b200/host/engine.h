// engine.h -- C++ host engine above the C ABI (include/cognn_b200.h): the device-resident counterpart of
// SSEdgeCentricAlgoKernel::onIteration / runAlgoKernelServer (include/ss_vertex_centric_algo_kernel.h:680-1189 of
// the reference) driving the CoGNN-Opt operators (algo_kernels/vertex_centric/optimize-gcn/gcn.h:198-887).
//
// One SSGcnEngine hosts one party (NcclComm: one process per GPU, the deployment shape) or all T parties
// (LoopbackComm: one process, one GPU -- the "both parties in one address space" driver used for parity tests).
// Shares never leave the device between operators; messages are flat little-endian u64 buffers.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/cognn_b200.h"

namespace cognn {

// GNNParam of the reference (include/task/task.h:78-170): same keys, same `key : value` file format.
struct GNNConfig {
    int num_layers = 2, num_labels = 0, input_dim = 0, hidden_dim = 0, num_samples = 0, num_edges = 0;
    double learning_rate = 0.5, train_ratio = 0.2, val_ratio = 0.2, test_ratio = 0.6;
    bool read(const std::string& file, std::string* err);
};

struct Message {
    uint64_t iter;
    int src, dst;
    std::string tag;
    std::vector<uint64_t> data;
};

// Inter-party message plane: replaces CommSync::send/recvShareVecVec (include/comm_sync.h:245-277) and the TaskComm
// channels.  Calls are posted, then completed together by exchange() (one NCCL group / one round of the protocol).
class Comm {
public:
    virtual ~Comm() {}
    virtual int world() const = 0;
    virtual bool is_local(int party) const = 0;
    virtual cgb_ctx* ctx() = 0;
    virtual void post_send(int src, int dst, const uint64_t* d, size_t n, const std::string& tag) = 0;
    virtual void post_recv(int dst, int src, uint64_t* d, size_t n) = 0;
    virtual void exchange() = 0;
    // all parties have reached this point and their streams are idle (no data message; used between the offline and the
    // online phase so that one party's dealer time is not counted as another party's online waiting time)
    virtual void barrier() {}
    // the engine announces, once its shapes are known, the most words one party sends another in one round; a plane may use it
    // to set up fixed receive arenas (NcclComm: peer-memory rounds instead of ncclSend / ncclRecv).  No data moves here.
    virtual void reserve(size_t /*max_words_per_pair_round*/) {}
    // name of the plane the rounds actually run on ("loopback", "nccl", "peer-memory (nccl bootstrap)")
    virtual const char* plane() const { return "loopback"; }
    bool record = false;        // keep a copy of every message sent by a local party (tests)
    uint64_t cur_iter = 0;
    std::vector<Message> transcript;
    uint64_t words_sent = 0;    // by local parties
    uint64_t rounds = 0;
};

std::unique_ptr<Comm> make_loopback_comm(int world, cgb_ctx* ctx);
// nccl_unique_id: 128 bytes from nccl_get_unique_id() on rank 0, distributed by the launcher (torch.distributed, MPI...)
std::unique_ptr<Comm> make_nccl_comm(int rank, int world, cgb_ctx* ctx, const void* nccl_unique_id);
void nccl_get_unique_id(void* out128);

// What a party derives from the reference's input files (graph_io_util.h:40-208) and preprocessing (ssk.h:295-534),
// flattened: local vertices ascending, degrees, and ONE CSR-by-destination over all its out-edges whose rows are the
// destination vertices of party 0, then party 1, ... (dummy self-edges of ssk.h:412-418 leave only their degree bump).
struct PartyGraph {
    int T = 0, me = 0;
    std::vector<uint64_t> vids;        // localVertexPos
    std::vector<uint32_t> offsets;     // T + 1: first output row of each destination party
    std::vector<uint64_t> in_deg_raw;  // before the dummy increment (feature normalisation, gcn.h:857-862)
    std::vector<uint64_t> in_deg;      // localVertexInDeg after the -r 1 dummy rule
    std::vector<uint32_t> rowptr, col;
    std::vector<uint8_t> is_border;    // isLocalVertexBorder (graph_io_util.h:169)
    cgb_csr* csr = nullptr;            // device ingest: the gather CSR is already resident (rowptr / col stay empty);
                                       // ownership passes to the engine in add_party
};
// edges: n_edges x 2 (src, dst) directed entries as in the .edge file; tid: vertex -> party (the .part file)
PartyGraph build_party_graph(const int64_t* edges, size_t n_edges, const int64_t* tid, size_t n_vertices, int T, int me);
// the same derivation by device passes (cgb_party_graph_build_host: scans, one edge pass, radix sort); the CSR never visits
// the host, only the per-vertex arrays come back.  This is what the engine's loaders use; the host builder above remains
// as the reference-shaped restatement the tests compare against.
PartyGraph build_party_graph_device(cgb_ctx* ctx, const int64_t* edges, size_t n_edges, const int64_t* tid, size_t n_vertices,
                                    int T, int me);

struct Metrics {
    uint64_t iter;
    int party;
    double loss, acc_full, acc_train, acc_test;
};

class SSGcnEngine {
public:
    SSGcnEngine(Comm* comm, const GNNConfig& cfg, int f, const uint32_t key[8]);
    ~SSGcnEngine();
    // raw (un-normalised) features n_local x F and labels of the party's vertices, in vids order
    void add_party(const PartyGraph& g, const double* feats_local, const int32_t* labels_local);
    void setup();                  // onAlgoKernelStart + vertex / weight sharing (gcn.h:854-887, ssk.h:200-232)
    void run(uint64_t n_iters);    // iterations [iter_, iter_ + n_iters)
    uint64_t iteration() const { return iter_; }
    // test / result access: name in {X, W0, W1, z0, z1, g, h_t0, h_t1}, role 0 = owner share, 1 = helper share of `owner`
    std::vector<uint64_t> download(int owner, int role, const std::string& name, uint32_t* rows, uint32_t* cols);
    const std::vector<Metrics>& metrics() const { return metrics_; }
    bool verbose = false;          // print the reference's log lines ("::iteration took", accuracy)
    double seconds_online = 0, seconds_offline = 0;
    double seconds_online_gpu = 0;  // the same online phases between two CUDA events on the main stream (no host launch / sync latency)
    double seconds_residual_host() const;   // 0: kept for the C ABI (the stand-ins moved to the device)
    // kernels that ran inside replayed CUDA graphs of the online phase (not seen by cgb_ctx_launch_count), and replays
    uint64_t replayed_launches() const;
    uint64_t eager_launches() const;  // kernels launched outside graph replays, over the main and all side streams
    uint64_t graph_replays() const;

    struct Impl;

private:
    Impl* impl_;
    uint64_t iter_ = 0;
    std::vector<Metrics> metrics_;
};

}  // namespace cognn
