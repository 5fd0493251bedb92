/*
 * cognn_b200.h -- C ABI of the B200-native share-local engine for CoGNN's secret-shared GCN path.
 *
 * This is the drop-in boundary ("level B" of SURVEY.md 8b): the entry points a CoGNN build binds in place of
 * the primitives its operators call but /root/reference does not define (Task-Worker / SCI-SilentOT / troy):
 *
 *   reference call site (file:line)                                     entry point here
 *   ------------------------------------------------------------------  -------------------------------------
 *   client/server_oblivious_mapper_online   ssk.h:752,760,818,848 /     cgb_expand_rows, cgb_sub (server mask)
 *                                           ssk.h:1011,1016,1057,1075
 *   prefix_network_aggregate(ADD_AGG)       optimize-gcn/gcn.h:328      cgb_segsum
 *   expand+ScatterComp+OGA+extract fused    ssk.h:751-821, gcn.h:300    cgb_csr_create + cgb_gather_sum
 *   sci::twoPartyGCNMatMul                  gcn.h:233,665,671,710       cgb_matmul, cgb_beaver_matmul_finish
 *   sci::twoPartyGCNVectorScale             gcn.h:247,476               cgb_rowmul_beaver_finish
 *   sci::twoPartyGCNCondVectorAddition      gcn.h:456                   cgb_rowmul_beaver_finish(f<0) + cgb_add,
 *                                                                       cgb_cond_add (selector local)
 *   sci::twoPartyGCNMatrixScale             gcn.h:676,723,764           cgb_scale_public
 *   sci::twoPartyGCNApplyGradient           gcn.h:678,730               cgb_apply_gradient
 *   sci::getPlainShareVecVec                gcn.h:604                   cgb_open_decode
 *   CryptoUtil::intoShares / encode / merge gcn.h:70,96,220,80          cgb_share_split, cgb_encode, cgb_decode
 *   transpose()                             gcn.h:230,648               cgb_transpose
 *   correlated randomness (OM masks, Beaver triples)                    cgb_prg_fill
 *   ("ssk.h" = include/ss_vertex_centric_algo_kernel.h, "gcn.h" = algo_kernels/vertex_centric/optimize-gcn/gcn.h)
 *
 * Conventions
 *   - Every tensor is a dense row-major array of uint64_t additive shares in Z_2^64 (no padding, ld == cols).
 *   - Pointers named d_* are DEVICE pointers on the context's GPU; h_* are HOST pointers.  No torch types.
 *   - `share` is 0 for the owner's share (sci::ALICE == 1 in the reference) and 1 for the helper's (sci::BOB == 2).
 *   - `f` is the number of fractional bits (SCALER_BIT_LENGTH in the reference, absent there; default CGB_SCALER_BITS).
 *   - All device entry points enqueue on the context's stream and return without synchronising unless stated.
 *   - Return value: 0 (CGB_OK) or a negative cgb_status; cgb_last_error() gives text.  The reference has no error
 *     returns (printf + exit(-1), ssk.h:794-797); the C++ shim in cognn_b200/host maps non-zero to that behaviour.
 *   - A context is single-threaded; distinct contexts are independent (one per (peer, role) thread of ssk.h:702-704).
 *   - There is NO CPU fallback: creating a context without a CUDA device fails with CGB_ERR_NO_DEVICE.
 */
#ifndef COGNN_B200_H_
#define COGNN_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGB_SCALER_BITS 16          /* default SCALER_BIT_LENGTH (must be < 31, gcn.h:191 uses int 1<<f) */
#define CGB_NO_ROW 0xFFFFFFFFu      /* expand_rows: position missing in source (allowMissing, ssk.h:848-851) */

typedef enum cgb_status {
    CGB_OK = 0,
    CGB_ERR_NO_DEVICE = -1,
    CGB_ERR_CUDA = -2,
    CGB_ERR_INVALID = -3,
    CGB_ERR_NOMEM = -4
} cgb_status;

typedef struct cgb_ctx cgb_ctx; /* device + stream + scratch */
typedef struct cgb_csr cgb_csr; /* device-resident CSR-by-destination + balanced work list; carries grow-only scratch for
                                   the gather, so one handle serves ONE stream at a time (create one per stream otherwise) */

/* ---- library / context ---------------------------------------------------------------------------------- */
const char* cgb_version(void);
int cgb_device_count(void);
int cgb_ctx_create(int device, cgb_ctx** out);
/* adopt an existing cudaStream_t (e.g. torch's current stream); stream == NULL -> legacy default stream */
int cgb_ctx_create_on_stream(int device, void* cuda_stream, cgb_ctx** out);
int cgb_ctx_destroy(cgb_ctx* ctx);
int cgb_ctx_sync(cgb_ctx* ctx);
void* cgb_ctx_stream(cgb_ctx* ctx);
const char* cgb_last_error(cgb_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t cgb_ctx_launch_count(cgb_ctx* ctx);

/* From now on every PRG launch of this context (cgb_prg_fill, cgb_prg_mask_sub, cgb_share_split) uses stream id
 * `stream + *d_bias`, the device word being read when the kernel RUNS: a CUDA graph captured once (one GAS iteration of
 * the engine) draws the randomness of a later iteration when it is replayed after the word was updated.  NULL switches
 * the bias off.  The word must stay allocated while such launches or graphs can still execute. */
int cgb_ctx_set_prg_stream_bias(cgb_ctx* ctx, const uint64_t* d_bias);

/* ---- memory --------------------------------------------------------------------------------------------- */
int cgb_malloc(cgb_ctx* ctx, size_t bytes, void** d_out);
int cgb_free(cgb_ctx* ctx, void* d_ptr);
int cgb_host_alloc(size_t bytes, void** h_out); /* pinned */
int cgb_host_free(void* h_ptr);
int cgb_memset(cgb_ctx* ctx, void* d_ptr, int value, size_t bytes);
int cgb_h2d(cgb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes); /* async on ctx stream */
int cgb_d2h(cgb_ctx* ctx, void* h_dst, const void* d_src, size_t bytes); /* async on ctx stream */
int cgb_d2d(cgb_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);

/* ---- (1) scatter / gather-sum ---------------------------------------------------------------------------- */
/* Builds the device CSR for one destination block from HOST arrays (index vectors are built once in
 * preprocessing, ssk.h:295-534): rowptr[n_rows+1], col[n_edges] (source row of each edge, rows grouped by
 * destination ascending = updateSrcVertexPos order).  Also builds the balanced work list used by
 * cgb_gather_sum.  Synchronises. */
int cgb_csr_create(cgb_ctx* ctx, const uint32_t* h_rowptr, const uint32_t* h_col, uint32_t n_rows, uint64_t n_edges,
                   uint32_t n_src_rows, cgb_csr** out);
/* same, from DEVICE arrays (copied; caller keeps ownership of its arrays) */
int cgb_csr_create_device(cgb_ctx* ctx, const uint32_t* d_rowptr, const uint32_t* d_col, uint32_t n_rows,
                          uint64_t n_edges, uint32_t n_src_rows, cgb_csr** out);
int cgb_csr_destroy(cgb_ctx* ctx, cgb_csr* csr);
uint64_t cgb_csr_num_edges(const cgb_csr* csr);
uint32_t cgb_csr_num_rows(const cgb_csr* csr);
const uint32_t* cgb_csr_rowptr(const cgb_csr* csr); /* device pointers */
const uint32_t* cgb_csr_col(const cgb_csr* csr);

/* y[v,:] = (d_delta ? d_delta[v,:] : 0) + sum_{e in row v} d_x[col[e],:]
 * d_x: n_src_rows x D;  d_delta, d_y: n_rows x D.  With d_x = x0 + (x1 - r) (own share plus the helper's OM online
 * message, one cgb_add) and d_delta = A*r - s (offline correlation) this is the whole client side of
 * expand -> ScatterComp -> prefix_network_aggregate -> extract in one pass over the edges.  d_y must not alias d_x. */
int cgb_gather_sum(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* d_x, const uint64_t* d_delta, uint64_t* d_y,
                   uint32_t D);
/* Same gather, but the output rows are split into n_blocks contiguous blocks (block t = rows [offsets[t], offsets[t+1]),
 * one per destination party) and block t is stored at d_block_base[t].  A base may be PEER memory -- another GPU's
 * receive buffer mapped with cgb_ipc_open -- so the mirror-update exchange (ssk.h:835 -> 1067/1090) is fused into the
 * gather: each row is written once, over NVLink, straight to the party that consumes it.  d_block_base and
 * block_row_offsets are HOST arrays (n_blocks <= 16).  Completion on the peer needs a cross-rank barrier afterwards. */
int cgb_gather_sum_blocks(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* d_x, const uint64_t* d_delta, uint32_t D,
                          uint32_t n_blocks, uint64_t* const* d_block_base, const uint32_t* block_row_offsets);
/* COMPACT form of the same gather: row k of d_y is the k-th destination row that HAS an edge (cgb_csr_nonempty_rows()[k],
 * ascending); rows without edges are not written.  d_y (and d_delta, if given) are cgb_csr_num_nonempty_rows() x D.  This is
 * the mirror-update block one party hands to another (ssk.h:835 -> 1067/1090): which destinations of the receiver have an
 * edge from the sender is public to both sides (sendPosVec / recvPosVec, ssk.h:507-516), so only those rows cross NVLink.
 * The consumer adds the block with cgb_scatter_add_rows. */
int cgb_gather_sum_compact(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* d_x, const uint64_t* d_delta, uint64_t* d_y,
                           uint32_t D);
/* The fused gather + exchange step: ONE launch over all of the party's out-edges whose output rows are cut into n_blocks
 * contiguous blocks (block b = rows [offsets[b], offsets[b+1]); the party's own block and the mirror-update blocks for the other
 * parties, or row pieces of them).  Block b is stored at d_block_base[b] -- densely (row - offsets[b]) or, if
 * block_compact[b], in the compact form of cgb_gather_sum_compact (position among the block's non-empty rows) -- and the moment
 * its last row has been stored, while the grid is still gathering the later blocks, the kernel raises d_block_flag[b]
 * (value flag_value, st.release.sys; the flag is usually in the CONSUMER's memory, mapped with cgb_ipc_open; NULL = no signal).
 * The consumer waits with cgb_flag_wait and pulls the block with cgb_scatter_add_rows, so the NVLink transfer of block b
 * overlaps the gather of blocks b+1... inside one kernel; this replaces the blocking sendShareVecVec / recvShareVecVec pairs of
 * ssk.h:1090-1100.  All arrays are HOST arrays (n_blocks <= 32).  Buffers of compact blocks hold
 * (number of non-empty rows of the block) x D words. */
int cgb_gather_sum_signal(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* d_x, uint32_t D, uint32_t n_blocks,
                          const uint32_t* block_row_offsets, uint64_t* const* d_block_base, const uint8_t* block_compact,
                          uint32_t* const* d_block_flag, uint32_t flag_value);
uint32_t cgb_csr_num_nonempty_rows(const cgb_csr* csr);
const uint32_t* cgb_csr_nonempty_rows(const cgb_csr* csr); /* device pointer, ascending row ids */
/* v[idx[k], :] (+)= src[k, :] for k < n: the GatherComp addition (gcn.h:456-463) of one received compact block.  d_src may
 * be PEER memory (the producer's staging buffer mapped with cgb_ipc_open): the rows are pulled over NVLink and added in one
 * pass, never landing in local memory first.  idx must not repeat (it is a PosVec: distinct destination rows); assign != 0
 * stores instead of adding.  n_ctas = 0 picks the grid. */
int cgb_scatter_add_rows(cgb_ctx* ctx, const uint32_t* d_idx, uint64_t n, const uint64_t* d_src, uint64_t* d_v, uint32_t D,
                         int assign, uint32_t n_ctas);
/* Cross-GPU arrival flags in place of the TCP recv that blocks ssk.h:1090: cgb_flag_signal stores `value` to a 4-byte flag
 * (usually in the consumer's memory, mapped with cgb_ipc_open) after everything enqueued before it on this context's stream is
 * visible system-wide; cgb_flag_wait holds this context's stream until *d_flag >= value (wrap-safe signed difference).
 * mode 0: stream memory operation (cuStreamWaitValue32, no SM occupied); mode 1: one-thread kernel polling for at most ~2 s,
 * then *d_err (optional device word) is set to 1 and the stream continues -- a lost peer cannot hang the GPU. */
int cgb_flag_signal(cgb_ctx* ctx, uint32_t* d_flag, uint32_t value);
int cgb_flag_wait(cgb_ctx* ctx, const uint32_t* d_flag, uint32_t value, int mode, uint32_t* d_err);
/* name of the kernel the last cgb_gather_sum* / cgb_matmul dispatch of this context chose (static string) */
const char* cgb_ctx_last_kernel(cgb_ctx* ctx);
/* CUDA IPC plumbing for the peer buffers (one process per GPU): d_ptr must come from cgb_malloc; handle is 64 bytes */
/* One protocol round between parties on different GPUs of a box without a library collective: the replacement of
 * CommSync::sendShareVecVec / recvShareVecVec (include/comm_sync.h:245-277) used by the engine's message plane.
 * A "link" is an ordered pair of parties; its receiver owns a double-buffered slot pair (2 * slot_words words, mapped by the
 * sender with cgb_ipc_open), a data flag (receiver's memory) and an acknowledge flag (sender's memory).
 * Both directions of a round go into ONE launch (a link says which it is); push segments must come first in `segs`, so a
 * rank's outgoing data is never queued behind CTAs that wait for incoming data.
 *   push link:     for every segment, wait until the peer acknowledged round s - 2 of the link, copy src -> dst + (s & 1) *
 *                  slot_words (dst = slot 0 address inside the PEER's arena); the last CTA of a link raises signal_flag to s.
 *   recv link:     wait for wait_flag >= s, copy src + (s & 1) * slot_words -> dst (src = slot 0 address inside the local
 *                  arena); the last CTA of a link raises signal_flag (the peer's acknowledge flag) to s.
 * s = *seq + 1 is read and advanced ON THE DEVICE, so the launches can be recorded into a CUDA graph.  seq, done and the
 * flags start at zero.  At most 16 segments and 16 links per call; waits are bounded (*d_err = 2 / 3 on a push / recv
 * timeout). */
typedef struct {
    const uint64_t* src;
    uint64_t* dst;
    uint64_t n_words;
    uint32_t link; /* index into the links array */
} cgb_xseg;
typedef struct {
    uint32_t* seq;             /* local: rounds completed on this link in this direction */
    uint32_t* done;            /* local: CTA completion counter, self-resetting */
    const uint32_t* wait_flag; /* local flag the peer writes (push: its acknowledgements, recv: its data flag) */
    uint32_t* signal_flag;     /* flag in the peer's memory (push: its data flag, recv: its acknowledge flag) */
    uint32_t recv;             /* 0: push link, 1: recv link */
} cgb_xlink;
int cgb_peer_round(cgb_ctx* ctx, const cgb_xseg* segs, uint32_t n_seg, const cgb_xlink* links, uint32_t n_links,
                   uint64_t slot_words, uint32_t ctas_per_seg, uint32_t* d_err);
int cgb_ipc_export(cgb_ctx* ctx, void* d_ptr, void* out_handle64);
int cgb_ipc_open(cgb_ctx* ctx, const void* handle64, void** d_peer_out);
int cgb_ipc_close(cgb_ctx* ctx, void* d_peer);
/* SM-driven bulk copy between two device buffers, either of which may be peer memory (cgb_ipc_open): the transport of the
 * mirror-update exchange where CommSync::sendShareVecVec / recvShareVecVec moved the block over TCP (comm_sync.h:245-277).
 * n_ctas CTAs of 128 threads (0 = 64); pointers and size must be multiples of 16 bytes. */
int cgb_peer_copy(cgb_ctx* ctx, void* d_dst, const void* d_src, size_t bytes, uint32_t n_ctas);
/* OM online, client side: y[j,:] = (idx[j]==CGB_NO_ROW ? 0 : x[idx[j],:]) + (delta ? delta[j,:] : 0) */
int cgb_expand_rows(cgb_ctx* ctx, const uint32_t* d_idx, uint64_t n_out, const uint64_t* d_x,
                    const uint64_t* d_delta, uint64_t* d_y, uint32_t D);
/* prefix_network_aggregate ADD_AGG over dst-sorted rows: segment s = rows [segptr[s], segptr[s+1]).
 * dup != 0: d_out is n_in x D and every row of a segment receives the segment sum; dup == 0: d_out is n_seg x D.
 * d_in and d_out must not alias. */
int cgb_segsum(cgb_ctx* ctx, const uint32_t* d_segptr, uint32_t n_seg, uint64_t n_in, const uint64_t* d_in,
               uint64_t* d_out, uint32_t D, int dup);

/* ---- graph ingest + index vectors on the device ------------------------------------------------ */
/* What a party derives from the reference's .edge and .part files before the first iteration: graphTilesFromEdgeList
 * (graph_io_util.h:40-208), GraphTile finalize (graph.h:607-641) and the index vectors of
 * SSEdgeCentricAlgoKernel (ss_vertex_centric_algo_kernel.h:295-534, with -r 1), flattened into
 *   vids        n_local      localVertexPos: the party's vertex ids, ascending
 *   offsets     T + 1 (host) first output row of each destination party (row = offsets[tid[d]] + index of d in its party)
 *   rowptr/col  ONE CSR-by-destination over all of the party's out-edges (rows = destinations of party 0, 1, ...;
 *               sources ascending inside a row, repeated edges kept): updateSrcVertexPos / updateDstVertexPos of every t
 *   in_deg_raw  in-degree over ALL edges (local, graph.h:627-632, and remote, graph_io_util.h:170-175)
 *   in_deg      the same after the dummy rule of ssk.h:412-418 (+1 for a vertex without a LOCAL in-edge)
 *   is_border   isLocalVertexBorder (graph_io_util.h:169)
 * d_edges: n_edges x 2 int64 (src, dst) as in the .edge file, 16-byte aligned; d_tid: n_vertices int64 (the .part file).
 * Everything is computed by device passes (scans, one edge pass, a radix sort); synchronises.  T <= 16. */
typedef struct cgb_party_graph cgb_party_graph;
int cgb_party_graph_build(cgb_ctx* ctx, const int64_t* d_edges, uint64_t n_edges, const int64_t* d_tid, uint64_t n_vertices,
                          int T, int me, cgb_party_graph** out);
/* same from HOST arrays (uploaded first) */
int cgb_party_graph_build_host(cgb_ctx* ctx, const int64_t* h_edges, uint64_t n_edges, const int64_t* h_tid,
                               uint64_t n_vertices, int T, int me, cgb_party_graph** out);
int cgb_party_graph_destroy(cgb_ctx* ctx, cgb_party_graph* g);
uint32_t cgb_party_graph_num_local(const cgb_party_graph* g);
uint32_t cgb_party_graph_num_rows(const cgb_party_graph* g);
uint64_t cgb_party_graph_num_out_edges(const cgb_party_graph* g);
const uint32_t* cgb_party_graph_offsets(const cgb_party_graph* g);    /* HOST, T + 1 */
const uint64_t* cgb_party_graph_vids(const cgb_party_graph* g);       /* device pointers from here on */
const uint64_t* cgb_party_graph_in_deg_raw(const cgb_party_graph* g);
const uint64_t* cgb_party_graph_in_deg(const cgb_party_graph* g);
const uint8_t* cgb_party_graph_is_border(const cgb_party_graph* g);
const uint32_t* cgb_party_graph_rowptr(const cgb_party_graph* g);
const uint32_t* cgb_party_graph_col(const cgb_party_graph* g);
/* the gather-sum CSR (with its balanced work list) of the party's out-edges, ready for cgb_gather_sum */
int cgb_party_graph_csr(cgb_ctx* ctx, const cgb_party_graph* g, cgb_csr** out);

/* ---- (2) dense contraction mod 2^64 ---------------------------------------------------------------------- */
/* C (M x N) = (accumulate ? C : 0) + op(A) * B.  op(A) = A (M x K), or A^T with A stored K x M if transA. */
int cgb_matmul(cgb_ctx* ctx, const uint64_t* d_A, const uint64_t* d_B, uint64_t* d_C, uint32_t M, uint32_t K,
               uint32_t N, int transA, int accumulate);
/* Beaver recombination: C_i = trunc_i( Z_i + E*V_i + U_i*F + [share==0] E*F ), E = X-U (M x K), F = W-V (K x N)
 * opened; f < 0 skips the truncation.  C must not alias the inputs. */
int cgb_beaver_matmul_finish(cgb_ctx* ctx, const uint64_t* d_E, const uint64_t* d_F, const uint64_t* d_U,
                             const uint64_t* d_V, const uint64_t* d_Z, uint64_t* d_C, uint32_t M, uint32_t K,
                             uint32_t N, int share, int f);
/* the same starting from the two halves of the opening: d_mine = this side's [E_i | F_i] (M*K + K*N words), d_peer the
 * peer's; d_mine is OPENED IN PLACE (mine += peer) by the launch that also forms V + F for share 0. */
int cgb_beaver_matmul_finish_open(cgb_ctx* ctx, uint64_t* d_mine, const uint64_t* d_peer, const uint64_t* d_U,
                                  const uint64_t* d_V, const uint64_t* d_Z, uint64_t* d_C, uint32_t M, uint32_t K,
                                  uint32_t N, int share, int f);

/* Pipe selection of cgb_matmul / cgb_beaver_matmul_finish for this context: -1 = CGB_MATMUL_IMPL from the environment or auto
 * (default), 0 = auto by shape, 1 = integer pipe (IMAD tiles), 2 = tensor pipe (tcgen05 int8 limbs).  Same bits either way. */
int cgb_ctx_set_matmul_impl(cgb_ctx* ctx, int impl);
/* Pipe ceilings measured on this device (synchronise; a few ms each): u64 multiply-adds per second of the integer pipe on
 * register operands, and u8 x u8 limb MACs per second of the tensor pipe on the limb kernel's own tcgen05.mma mix with
 * operands already in shared memory.  bench.py divides the matmul rates by these. */
int cgb_probe_imad_peak(cgb_ctx* ctx, double* u64_mac_per_s);
int cgb_probe_tensor_i8_peak(cgb_ctx* ctx, double* limb_mac_per_s);

/* ---- (3) fused elementwise: truncation, masking / opening, scaling ---------------------------------------- */
int cgb_add(cgb_ctx* ctx, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, uint64_t n);
int cgb_sub(cgb_ctx* ctx, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, uint64_t n);
/* out = sum of n_in (<= 16) share vectors: the GatherComp additions over all source parties (gcn.h:456-463, called T
 * times per iteration in the reference) in one pass.  d_in is a HOST array of device pointers; out may alias d_in[0]. */
int cgb_sum_n(cgb_ctx* ctx, const uint64_t* const* d_in, uint32_t n_in, uint64_t* d_out, uint64_t n);
int cgb_trunc(cgb_ctx* ctx, const uint64_t* d_x, uint64_t* d_out, uint64_t n, int f, int share);
int cgb_scale_public(cgb_ctx* ctx, const uint64_t* d_x, uint64_t c, uint64_t* d_out, uint64_t n, int f, int share);
int cgb_apply_gradient(cgb_ctx* ctx, const uint64_t* d_W, const uint64_t* d_d, uint64_t lr, uint64_t* d_out,
                       uint64_t n, int f, int share);
/* out = trunc_i( c + e*b[row] + fv[row]*a + [share==0] e*fv[row] ), f < 0: no truncation. */
int cgb_rowmul_beaver_finish(cgb_ctx* ctx, const uint64_t* d_e, const uint64_t* d_fv, const uint64_t* d_a,
                             const uint64_t* d_b, const uint64_t* d_c, uint64_t* d_out, uint64_t rows, uint32_t D,
                             int share, int f);
/* weight-gradient step in one pass (gcn.h:673-678): d' = trunc_i(d * gs), W' = W - trunc_i(d' * lr); d_d_out may be NULL
 * or alias d_d, d_W_out may alias d_W. */
int cgb_scale_apply_gradient(cgb_ctx* ctx, const uint64_t* d_W, const uint64_t* d_d, uint64_t gs, uint64_t lr,
                             uint64_t* d_d_out, uint64_t* d_W_out, uint64_t n, int f, int share);
/* weight averaging (gcn.h:747-777): every d_out[k] = trunc_i( (sum_j d_in[j]) * c ); HOST arrays of device pointers
 * (<= 16 inputs, <= 4 outputs); outputs may alias inputs. */
int cgb_avg_public(cgb_ctx* ctx, const uint64_t* const* d_in, uint32_t n_in, uint64_t c, uint64_t* const* d_out,
                   uint32_t n_out, uint64_t n, int f, int share);
/* the same with the opening folded in: e = mine + peer over the message layout [ rows x D words | rows scaler words ] that
 * sci::twoPartyGCNVectorScale's parties exchange (gcn.h:247,476); saves the separate pass that opens the message. */
int cgb_rowmul_beaver_finish_open(cgb_ctx* ctx, const uint64_t* d_mine, const uint64_t* d_peer, const uint64_t* d_a,
                                  const uint64_t* d_b, const uint64_t* d_c, uint64_t* d_out, uint64_t rows, uint32_t D,
                                  int share, int f);
/* out = a * b[row] - c (rows x D; b has one word per row): the dealer's third row-scaling share c1 = (a0+a1)(b0+b1)[row] - c0. */
int cgb_rowmul_sub(cgb_ctx* ctx, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_c, uint64_t* d_out,
                   uint64_t rows, uint32_t D);
/* one launch for the two halves of a Beaver message: out[0,n0) = a0 - b0, out[n0,n0+n1) = a1 - b1 (a1 == NULL: 0 - b1),
 * i.e. [X - U | W - V] of twoPartyGCNMatMul (gcn.h:233) or [x - a | s - b] of twoPartyGCNVectorScale (gcn.h:247). */
int cgb_sub_pair(cgb_ctx* ctx, const uint64_t* d_a0, const uint64_t* d_b0, uint64_t n0, const uint64_t* d_a1,
                 const uint64_t* d_b1, uint64_t n1, uint64_t* d_out);
int cgb_cond_add(cgb_ctx* ctx, const uint64_t* d_v, const uint64_t* d_u, const uint8_t* d_cond, uint64_t* d_out,
                 uint64_t rows, uint32_t D);
/* n_seg independent device-to-device copies of n_words[j] words in one launch per 16 segments (HOST arrays of device
 * pointers; segments must not overlap): all messages of one communication round between parties hosted on one GPU
 * (the single-process form of CommSync::sendShareVecVec / recvShareVecVec, include/comm_sync.h:245-277). */
int cgb_copy_segments(cgb_ctx* ctx, uint64_t* const* d_dst, const uint64_t* const* d_src, const uint64_t* n_words,
                      uint32_t n_seg);
int cgb_transpose(cgb_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint32_t rows, uint32_t cols);
int cgb_encode(cgb_ctx* ctx, const double* d_x, uint64_t* d_out, uint64_t n, int f);
int cgb_decode(cgb_ctx* ctx, const uint64_t* d_v, double* d_out, uint64_t n, int f);
/* s1 = PRG(key, stream, word_offset ...), s0 = enc(x) - s1 */
int cgb_share_split(cgb_ctx* ctx, const double* d_x, uint64_t n, int f, const uint32_t key[8], uint64_t stream,
                    uint64_t word_offset, uint64_t* d_s0, uint64_t* d_s1);
int cgb_open_decode(cgb_ctx* ctx, const uint64_t* d_s0, const uint64_t* d_s1, double* d_out, uint64_t n, int f);
/* Loss / accuracy the owner prints after the prediction layer (gcn.h:603-632), from the two shares of the n x C probabilities:
 * every 256-vertex block writes 4 doubles {sum of -log p[label], hits over all / the first train_rows / the rows from
 * train_rows + val_rows on} to d_block_out (zeros count as 0.001, gcn.h:615; first maximal class wins); the caller adds the
 * *n_blocks records in order.  d_block_out == NULL only reports *n_blocks. */
int cgb_prediction_metrics(cgb_ctx* ctx, const uint64_t* d_s0, const uint64_t* d_s1, const int32_t* d_labels, uint32_t n,
                           uint32_t C, uint32_t train_rows, uint32_t val_rows, int f, double* d_block_out,
                           uint32_t* n_blocks);

/* ---- (4) device PRG (ChaCha20 block function, RFC 8439; stream layout in DESIGN.md) ------------------------ */
int cgb_prg_fill(cgb_ctx* ctx, const uint32_t key[8], uint64_t stream, uint64_t word_offset, uint64_t* d_out,
                 uint64_t n_words);
/* out = in - PRG(...): the server side of the OM online message (x1 - r) without materialising r */
int cgb_prg_mask_sub(cgb_ctx* ctx, const uint32_t key[8], uint64_t stream, uint64_t word_offset,
                     const uint64_t* d_in, uint64_t* d_out, uint64_t n_words);
/* Several fills in one launch: out = PRG(stream_a) [+ PRG(stream_b) if has_b], and out_b (optional) = PRG(stream_b) alone --
 * what the dealer emulation hands one side for one Beaver triple or OM correlation (word offset 0). */
typedef struct {
    uint64_t* out;
    uint64_t* out_b;
    uint64_t n_words;
    uint64_t stream_a, stream_b;
    uint32_t has_b;
} cgb_prg_seg;
int cgb_prg_fill_multi(cgb_ctx* ctx, const uint32_t key[8], const cgb_prg_seg* segs, uint32_t n_seg);
/* out = sum_j d_in[j] + sum_k PRG(key, streams[k], 0 ...): the GatherComp additions of one destination party
 * (gcn.h:456-463) with the OM mask shares regenerated in registers.  `streams` and `d_in` are HOST arrays (<= 16 each);
 * out may alias an input. */
int cgb_prg_sum(cgb_ctx* ctx, const uint32_t key[8], const uint64_t* streams, uint32_t n_streams,
                const uint64_t* const* d_in, uint32_t n_in, uint64_t* d_out, uint64_t n_words);

/* ---- 2PC-RESIDUAL stand-ins (IDEAL FUNCTIONALITY, NOT SECURE) ------------------------------------------------- */
/* sci::twoPartyGCNRelu (gcn.h:549) and the ReLU' mask of sci::twoPartyGCNBackwardNNWithoutAH (gcn.h:705) stay on the
 * reference's MPC backend.  So that an epoch can run end to end, the engine evaluates them on the RECONSTRUCTED value
 * after the helper has sent its share to the owner -- a stand-in for the 2PC, exact integer arithmetic:
 *   relu:      out = (int64)(a0+a1) > 0 ? a0+a1 : 0
 *   relu_grad: out = (int64)(z0+z1) > 0 ? g0+g1 : 0                                                              */
/* the stand-in and the re-sharing of its result in one pass: out = relu-or-gate(...) - PRG(key, stream, 0 ...), i.e. the
 * owner's new share when the helper's is the PRG stream; d_z0 == d_z1 == NULL: ReLU of a, else ReLU'(z) applied to a. */
int cgb_ideal_relu_reshare(cgb_ctx* ctx, const uint32_t key[8], uint64_t stream, const uint64_t* d_a0,
                           const uint64_t* d_a1, const uint64_t* d_z0, const uint64_t* d_z1, uint64_t* d_out,
                           uint64_t n_words);
int cgb_ideal_relu(cgb_ctx* ctx, const uint64_t* d_a0, const uint64_t* d_a1, uint64_t* d_out, uint64_t n);
int cgb_ideal_relu_grad(cgb_ctx* ctx, const uint64_t* d_g0, const uint64_t* d_g1, const uint64_t* d_z0,
                        const uint64_t* d_z1, uint64_t* d_out, uint64_t n);
/* sci::twoPartyGCNForwardNNPredictionWithoutWeight (gcn.h:578,591), same kind of stand-in: row-wise softmax of the
 * reconstructed logits z0+z1 (n x C) in double with exp restated from IEEE + - * only (bit-identical to the oracle's
 * orc_det_exp), P = enc(p), pmy = P - (onehot(label) << f) on the first train_rows rows and 0 below (gcn.h:639-641).
 * d_labels: n int32 class ids on the device. */
int cgb_ideal_softmax(cgb_ctx* ctx, const uint64_t* d_z0, const uint64_t* d_z1, const int32_t* d_labels, uint64_t n,
                      uint32_t C, uint64_t train_rows, int f, uint64_t* d_P, uint64_t* d_pmy);

/* ---- host-buffer entry point (what a CoGNN operator holding std::vector data calls) ------------------------ */
/* One gather-sum step with HOST share rows: H2D of x (and delta if given), kernel, D2H of y,
 * synchronises.  h_x / h_y should be pinned (cgb_host_alloc) for full PCIe rate.  The CSR stays device resident.  bench.py's `e2e` times this. */
int cgb_host_gather_sum(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* h_x, const uint64_t* h_delta,
                        uint64_t* h_y, uint32_t D);

/* Pipelined variant for back-to-back steps: returns after enqueueing; H2D of step i+1 overlaps the kernel of step i and
 * the D2H of step i-1 (two staging slots, three streams).  Host buffers must be pinned and stay valid until
 * cgb_host_sync() returns.  Results are identical to cgb_host_gather_sum. */
int cgb_host_gather_sum_async(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* h_x, const uint64_t* h_delta,
                              uint64_t* h_y, uint32_t D);
int cgb_host_sync(cgb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* COGNN_B200_H_ */
