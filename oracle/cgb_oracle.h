/*
 * cgb_oracle.h -- CPU ORACLE for the share-local arithmetic of CoGNN's secret-shared GCN path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is shipped or measured as the product; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the arithmetic this restates (sci::twoPartyGCN*, prefix_network_aggregate,
 * *_oblivious_mapper_online, CryptoUtil::*) is only CALLED in /root/reference, never defined there
 * (it lives in the un-vendored Task-Worker / SCI-SilentOT / troy trees, no version pins), and the
 * reference tree holds no tests or golden vectors.  What IS pinned:
 *   - the PRG against RFC 8439 ChaCha20 test vectors and the `cryptography` package (OpenSSL),
 *   - the index vectors against the worked example derived from ss_vertex_centric_algo_kernel.h:295-534,
 *   - weight init against glibc rand() (optimize-gcn/gcn.h:838-852),
 *   - all ring arithmetic against Python big-int arithmetic,
 *   - the reconstructed epoch against a float64 GCN of the same dataflow.
 * Every unpinned decision is a named constant documented in DESIGN.md "Frozen semantics".
 *
 * All tensors are dense row-major uint64_t (additive shares in Z_2^64), no padding.
 */
#ifndef CGB_ORACLE_H_
#define CGB_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NO_ROW 0xFFFFFFFFu /* expand_rows: "allowMissing" position (ss_vertex_centric_algo_kernel.h:848-851) */

/* ---- (4) PRG: ChaCha20 block function, RFC 8439 section 2.3 ------------------------------------ */
void orc_chacha20_block(const uint32_t key[8], uint32_t counter, const uint32_t nonce[3], uint32_t out[16]);
/* Word w (u64, little endian) of stream `stream` lives in block b = w/8 at byte offset 8*(w%8);
 * block b uses counter = (uint32_t)b and nonce = { (u32)stream, (u32)(stream>>32), (u32)(b>>32) }. */
void orc_prg_fill(const uint32_t key[8], uint64_t stream, uint64_t word_offset, uint64_t* out, size_t n_words);

/* ---- (3) fixed point, share split / open, truncation, public scaling ---------------------------- */
/* CryptoUtil::encodeDoubleAsFixedPoint call sites optimize-gcn/gcn.h:220,473,538 and the public-scalar
 * pattern static_cast<uint64_t>(x * (1<<SCALER_BIT_LENGTH)) at gcn.h:191,676,678,764: C truncation toward 0. */
uint64_t orc_encode_fixed(double x, int f);
double orc_decode_fixed(uint64_t v, int f);
void orc_encode(const double* x, uint64_t* out, size_t n, int f);
void orc_decode(const uint64_t* v, double* out, size_t n, int f);
/* CryptoUtil::intoShares (gcn.h:70,96): s1 = PRG word, s0 = enc(x) - s1 (owner keeps s0). */
void orc_share_split(const double* x, size_t n, int f, const uint32_t key[8], uint64_t stream,
                     uint64_t word_offset, uint64_t* s0, uint64_t* s1);
/* CryptoUtil::mergeShareAsDouble (gcn.h:80) / sci::getPlainShareVecVec (gcn.h:604). */
void orc_open_decode(const uint64_t* s0, const uint64_t* s1, double* out, size_t n, int f);
void orc_add(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
void orc_sub(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* SecureML local truncation: share 0: z >> f (logical); share 1: -((-z) >> f). */
uint64_t orc_trunc_share(uint64_t z, int f, int share);
void orc_trunc(const uint64_t* x, uint64_t* out, size_t n, int f, int share);
/* sci::twoPartyGCNMatrixScale (gcn.h:676,723,764): out = trunc(x * c), c public. */
void orc_scale_public(const uint64_t* x, uint64_t c, uint64_t* out, size_t n, int f, int share);
/* sci::twoPartyGCNApplyGradient (gcn.h:678,730): out = W - trunc(d * lr). */
void orc_apply_gradient(const uint64_t* W, const uint64_t* d, uint64_t lr, uint64_t* out, size_t n, int f, int share);
/* Share-local finish of an elementwise Beaver product x (rows x D, shared) times s (one per row, shared):
 * out_i = c_i + e*b_i + fv*a_i + [share==0] e*fv, then truncated by f bits if f >= 0.
 * Serves sci::twoPartyGCNVectorScale (gcn.h:247,476) with f = SCALER bits and the MUX of
 * sci::twoPartyGCNCondVectorAddition (gcn.h:456) with f < 0 (selector bit is an integer). */
void orc_rowmul_beaver_finish(const uint64_t* e, const uint64_t* fv, const uint64_t* a, const uint64_t* b,
                              const uint64_t* c, uint64_t* out, size_t rows, size_t D, int share, int f);
/* out = v + (cond[row] ? u : 0) with cond public to the caller: the share-local add of gcn.h:456-463. */
void orc_cond_add(const uint64_t* v, const uint64_t* u, const uint8_t* cond, uint64_t* out, size_t rows, size_t D);
void orc_transpose(const uint64_t* in, uint64_t* out, size_t rows, size_t cols); /* task.h transpose(), gcn.h:230,648 */

/* ---- 2PC-RESIDUAL stand-ins (ideal functionality on RECONSTRUCTED values; not secure, not the reference's protocol) ---- */
/* exp() restated with IEEE-754 double +, -, * only (no libm, no fused multiply-add), so that the C oracle, the numpy
 * restatement and the CUDA kernel produce the same bits:  k = floor(x*log2(e) + 1/2),  r = (x - k*ln2_hi) - k*ln2_lo,
 * exp(r) = sum_{i<=13} r^i / i! by Horner, result = that times 2^k.  x < -700 gives 0, x > 700 is clamped. */
double orc_det_exp(double x);
/* sci::twoPartyGCNForwardNNPredictionWithoutWeight (gcn.h:578,591) on the reconstructed logits z = z0 + z1 (n x C):
 * row-wise softmax in double (max subtracted, left-to-right sum), P = enc(p), pmy = P - onehot(label)<<f, rows >= train_rows
 * get pmy = 0 (gcn.h:639-641). */
void orc_ideal_softmax(const uint64_t* z0, const uint64_t* z1, const int32_t* labels, size_t n, size_t C,
                       size_t train_rows, int f, uint64_t* P, uint64_t* pmy);

/* ---- (1) scatter / gather-sum ------------------------------------------------------------------- */
/* Fused expand + group-by-destination sum: y[v] = (delta ? delta[v] : 0) + sum_{e in [rowptr[v],rowptr[v+1])} x[col[e]].
 * Composite of the OM expand (ss_vertex_centric_algo_kernel.h:751-763), ScatterComp copy (gcn.h:300),
 * prefix_network_aggregate ADD_AGG (gcn.h:328-335) and the extract OM (ssk.h:818-821). */
void orc_gather_sum_csr(const uint32_t* rowptr, const uint32_t* col, const uint64_t* x, const uint64_t* delta,
                        uint64_t* y, size_t n_rows, size_t D);
/* OM online, client side: y[j] = (idx[j]==ORC_NO_ROW ? 0 : x[idx[j]]) + (delta ? delta[j] : 0). */
void orc_expand_rows(const uint32_t* idx, size_t n_out, const uint64_t* x, const uint64_t* delta, uint64_t* y, size_t D);
/* prefix_network_aggregate(dstPos, svv, ADD_AGG, ...) on a dst-sorted E x D block: segment s covers rows
 * [segptr[s], segptr[s+1]).  dup != 0: E x D output, every row of a segment holds the segment sum ("duplicated"
 * layout kept by UpdatePreMergeComp, gcn.h:328);  dup == 0: n_seg x D output. */
void orc_segsum(const uint32_t* segptr, size_t n_seg, const uint64_t* in, uint64_t* out, size_t D, int dup);

/* ---- (2) dense contraction mod 2^64 ------------------------------------------------------------- */
/* C (M x N) = [accumulate ? C : 0] + op(A) * B;  op(A) = A (M x K) or, if transA, A^T with A stored K x M. */
void orc_matmul(const uint64_t* A, const uint64_t* B, uint64_t* C, size_t M, size_t K, size_t N, int transA, int accumulate);
/* Beaver recombination for sci::twoPartyGCNMatMul (gcn.h:233,665,671,710):
 * C_i = trunc_i( Z_i + E*V_i + U_i*F + [share==0] E*F ),  E = X - U and F = W - V opened.  f < 0: no truncation. */
void orc_beaver_matmul_finish(const uint64_t* E, const uint64_t* F, const uint64_t* U, const uint64_t* V,
                              const uint64_t* Z, uint64_t* C, size_t M, size_t K, size_t N, int share, int f);

/* threads used by the OpenMP loops (1 if built without OpenMP) */
int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif /* CGB_ORACLE_H_ */
