/* mock_cgb.c -- TEST INFRASTRUCTURE ONLY.  A host-memory stand-in for the subset of the C ABI (include/cognn_b200.h) that the
 * reference-API shim calls, implemented with the CPU oracle (oracle/cgb_oracle.c).  It exists so that the CPU test suite
 * (-m "not gpu") can exercise the HOST LOGIC of the drop-in -- the reference's own harness.cpp / ss_vertex_centric_algo_kernel.h /
 * gcn.h compiled against cognn_b200/host/shim, its transport, its call sequence -- in a container without a GPU.  It is built
 * into tests/mock/libcognn_b200.so and selected with LD_LIBRARY_PATH by tests/test_reference_dropin.py only; nothing in the
 * product, bench.py or the GPU tests loads it (those use cognn_b200/libcognn_b200.so, which has no CPU path). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../oracle/cgb_oracle.h"

typedef struct cgb_ctx { int dummy; } cgb_ctx;

int cgb_ctx_create(int device, cgb_ctx** out) { (void)device; *out = (cgb_ctx*)calloc(1, sizeof(cgb_ctx)); return 0; }
int cgb_ctx_destroy(cgb_ctx* c) { free(c); return 0; }
int cgb_ctx_sync(cgb_ctx* c) { (void)c; return 0; }
const char* cgb_last_error(cgb_ctx* c) { (void)c; return "mock"; }
int cgb_malloc(cgb_ctx* c, size_t bytes, void** out) { (void)c; *out = malloc(bytes ? bytes : 8); return *out ? 0 : -4; }
int cgb_free(cgb_ctx* c, void* p) { (void)c; free(p); return 0; }
int cgb_memset(cgb_ctx* c, void* p, int v, size_t n) { (void)c; memset(p, v, n); return 0; }
int cgb_h2d(cgb_ctx* c, void* d, const void* h, size_t n) { (void)c; memcpy(d, h, n); return 0; }
int cgb_d2h(cgb_ctx* c, void* h, const void* d, size_t n) { (void)c; memcpy(h, d, n); return 0; }
int cgb_add(cgb_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* o, uint64_t n) { (void)c; orc_add(a, b, o, n); return 0; }
int cgb_sub(cgb_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* o, uint64_t n) { (void)c; orc_sub(a, b, o, n); return 0; }
int cgb_prg_fill(cgb_ctx* c, const uint32_t key[8], uint64_t stream, uint64_t off, uint64_t* out, uint64_t n) {
    (void)c; orc_prg_fill(key, stream, off, out, n); return 0;
}
int cgb_prg_mask_sub(cgb_ctx* c, const uint32_t key[8], uint64_t stream, uint64_t off, const uint64_t* in, uint64_t* out, uint64_t n) {
    (void)c;
    uint64_t* r = (uint64_t*)malloc((n ? n : 1) * 8);
    orc_prg_fill(key, stream, off, r, n);
    for (uint64_t i = 0; i < n; ++i) out[i] = in[i] - r[i];
    free(r);
    return 0;
}
int cgb_matmul(cgb_ctx* c, const uint64_t* A, const uint64_t* B, uint64_t* C, uint32_t M, uint32_t K, uint32_t N, int tA, int acc) {
    (void)c; orc_matmul(A, B, C, M, K, N, tA, acc); return 0;
}
int cgb_beaver_matmul_finish(cgb_ctx* c, const uint64_t* E, const uint64_t* F, const uint64_t* U, const uint64_t* V, const uint64_t* Z,
                             uint64_t* C, uint32_t M, uint32_t K, uint32_t N, int share, int f) {
    (void)c; orc_beaver_matmul_finish(E, F, U, V, Z, C, M, K, N, share, f); return 0;
}
int cgb_rowmul_beaver_finish(cgb_ctx* c, const uint64_t* e, const uint64_t* fv, const uint64_t* a, const uint64_t* b, const uint64_t* cc,
                             uint64_t* out, uint64_t rows, uint32_t D, int share, int f) {
    (void)c; orc_rowmul_beaver_finish(e, fv, a, b, cc, out, rows, D, share, f); return 0;
}
int cgb_scale_public(cgb_ctx* c, const uint64_t* x, uint64_t k, uint64_t* out, uint64_t n, int f, int share) {
    (void)c; orc_scale_public(x, k, out, n, f, share); return 0;
}
int cgb_apply_gradient(cgb_ctx* c, const uint64_t* W, const uint64_t* d, uint64_t lr, uint64_t* out, uint64_t n, int f, int share) {
    (void)c; orc_apply_gradient(W, d, lr, out, n, f, share); return 0;
}
int cgb_open_decode(cgb_ctx* c, const uint64_t* s0, const uint64_t* s1, double* out, uint64_t n, int f) {
    (void)c; orc_open_decode(s0, s1, out, n, f); return 0;
}
int cgb_expand_rows(cgb_ctx* c, const uint32_t* idx, uint64_t n_out, const uint64_t* x, const uint64_t* delta, uint64_t* y, uint32_t D) {
    (void)c; orc_expand_rows(idx, n_out, x, delta, y, D); return 0;
}
int cgb_segsum(cgb_ctx* c, const uint32_t* segptr, uint32_t n_seg, uint64_t n_in, const uint64_t* in, uint64_t* out, uint32_t D, int dup) {
    (void)c; (void)n_in; orc_segsum(segptr, n_seg, in, out, D, dup); return 0;
}
int cgb_ideal_relu(cgb_ctx* c, const uint64_t* a0, const uint64_t* a1, uint64_t* out, uint64_t n) {
    (void)c;
    for (uint64_t i = 0; i < n; ++i) { uint64_t v = a0[i] + a1[i]; out[i] = (int64_t)v > 0 ? v : 0; }
    return 0;
}
int cgb_ideal_relu_grad(cgb_ctx* c, const uint64_t* g0, const uint64_t* g1, const uint64_t* z0, const uint64_t* z1, uint64_t* out, uint64_t n) {
    (void)c;
    for (uint64_t i = 0; i < n; ++i) out[i] = (int64_t)(z0[i] + z1[i]) > 0 ? g0[i] + g1[i] : 0;
    return 0;
}
int cgb_ideal_softmax(cgb_ctx* c, const uint64_t* z0, const uint64_t* z1, const int32_t* labels, uint64_t n, uint32_t C, uint64_t train_rows,
                      int f, uint64_t* P, uint64_t* pmy) {
    (void)c; orc_ideal_softmax(z0, z1, labels, n, C, train_rows, f, P, pmy); return 0;
}
