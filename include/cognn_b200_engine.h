/*
 * cognn_b200_engine.h -- C ABI of the host engine (cognn_b200/host): the device-resident counterpart of the reference's
 * SSEdgeCentricAlgoKernel iteration driver (include/ss_vertex_centric_algo_kernel.h:167-277, 680-1189) running the
 * CoGNN-Opt GCN operators (algo_kernels/vertex_centric/optimize-gcn/gcn.h).  One engine hosts one party (NCCL plane,
 * one process per GPU; replaces include/engine.h:143-222 + include/comm_sync.h) or all parties (loopback plane).
 * Plain pointers and sizes only.  Return 0 on success, -1 on error (cge_last_error).
 */
#ifndef COGNN_B200_ENGINE_H_
#define COGNN_B200_ENGINE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cge_engine cge_engine;

/* GNNParam of the reference (include/task/task.h:88-97) + engine settings */
typedef struct cge_config {
    int32_t num_layers, num_labels, input_dim, hidden_dim, num_samples, num_edges;
    double learning_rate, train_ratio, val_ratio, test_ratio;
    int32_t scaler_bits;       /* SCALER_BIT_LENGTH */
    uint32_t key[8];           /* dealer PRG key */
    int32_t record_messages;   /* keep a copy of every message a local party sends (tests) */
    int32_t verbose;           /* print the reference's log lines */
} cge_config;

const char* cge_last_error(cge_engine* h);
int cge_nccl_unique_id(void* out128);
int cge_create_loopback(int device, void* cuda_stream, int n_parties, const cge_config* cfg, cge_engine** out);
int cge_create_nccl(int device, void* cuda_stream, int rank, int n_parties, const void* nccl_uid128, const cge_config* cfg,
                    cge_engine** out);
int cge_destroy(cge_engine* h);
/* edges: n_edges x 2 (src, dst) as in the reference's .edge file; tid: vertex -> party (.part file);
 * feats_global: n_vertices x input_dim raw features, labels_global: n_vertices (.vertex file) */
int cge_add_party(cge_engine* h, int party, const int64_t* edges, uint64_t n_edges, const int64_t* tid, uint64_t n_vertices,
                  const double* feats_global, const int32_t* labels_global);
int cge_setup(cge_engine* h);
int cge_run(cge_engine* h, uint64_t n_iters);
/* name in {X, W0, W1, z0, z1, g, h_t0, h_t1, V, Xp}; role 0 = owner's share, 1 = helper's share of `owner`.
 * Returns the element count (out may be NULL to query), -1 on error. */
int64_t cge_download(cge_engine* h, int owner, int role, const char* name, uint64_t* out, uint64_t capacity, uint32_t* rows,
                     uint32_t* cols);
uint64_t cge_message_count(cge_engine* h);
int cge_message_info(cge_engine* h, uint64_t i, uint64_t* iter, int* src, int* dst, char* tag, uint64_t tag_cap, uint64_t* n_words);
int cge_message_data(cge_engine* h, uint64_t i, uint64_t* out, uint64_t capacity);
uint64_t cge_words_sent(cge_engine* h);
uint64_t cge_rounds(cge_engine* h);
/* the message plane the protocol rounds run on: "loopback", "nccl", or "peer-memory rounds over NVLink (...)" */
const char* cge_plane(cge_engine* h);
uint64_t cge_launch_count(cge_engine* h);   /* kernels launched, including those inside replayed CUDA graphs */
/* From the second epoch on, the online phase of each GAS iteration is one CUDA graph (captured at its first later
 * occurrence, replayed afterwards; COGNN_B200_GRAPHS=0 or record_messages keep the eager path).  Number of replays: */
uint64_t cge_graph_replays(cge_engine* h);
double cge_seconds_online(cge_engine* h);   /* host wall time of the online phases, stream-synchronised */
double cge_seconds_online_gpu(cge_engine* h); /* the same phases between two CUDA events on the engine's main stream */
double cge_seconds_offline(cge_engine* h);  /* dealer emulation (correlation generation), excluded from online */
double cge_seconds_residual_host(cge_engine* h); /* always 0: the 2PC-residual stand-ins run on the device (cgb_ideal_*) */
uint64_t cge_metrics_count(cge_engine* h);
int cge_metrics_get(cge_engine* h, uint64_t i, uint64_t* iter, int* party, double* loss, double* acc_full, double* acc_train,
                    double* acc_test);
/* index-vector builder alone (ssk.h:295-534 with -r 1), host only: first call with NULL arrays to get the sizes */
int cge_build_party_graph(const int64_t* edges, uint64_t n_edges, const int64_t* tid, uint64_t n_vertices, int T, int me,
                          uint64_t* vids, uint64_t* in_deg_raw, uint64_t* in_deg, uint32_t* offsets, uint32_t* rowptr,
                          uint32_t* col, uint64_t* n_local, uint64_t* n_rows, uint64_t* n_col);

#ifdef __cplusplus
}
#endif
#endif
