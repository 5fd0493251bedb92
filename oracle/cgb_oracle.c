/*
 * cgb_oracle.c -- CPU ORACLE (test infrastructure, see cgb_oracle.h).  Plain C, OpenMP on the outer loops
 * so the same code doubles as the "port" CPU baseline in bench.py.  Parity unpinned by the reference
 * (SURVEY.md 8c); pinned against RFC 8439, OpenSSL, glibc rand and Python big-ints in tests/.
 */
#include "cgb_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------ */
/* (4) PRG: RFC 8439 section 2.3 ChaCha20 block function                                              */
/* ------------------------------------------------------------------------------------------------ */
static inline uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
#define ORC_QR(a, b, c, d)   \
    a += b; d ^= a; d = rotl32(d, 16); \
    c += d; b ^= c; b = rotl32(b, 12); \
    a += b; d ^= a; d = rotl32(d, 8);  \
    c += d; b ^= c; b = rotl32(b, 7);

void orc_chacha20_block(const uint32_t key[8], uint32_t counter, const uint32_t nonce[3], uint32_t out[16]) {
    uint32_t s[16], x[16];
    s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
    for (int i = 0; i < 8; ++i) s[4 + i] = key[i];
    s[12] = counter;
    s[13] = nonce[0]; s[14] = nonce[1]; s[15] = nonce[2];
    memcpy(x, s, sizeof(x));
    for (int r = 0; r < 10; ++r) {
        ORC_QR(x[0], x[4], x[8], x[12])
        ORC_QR(x[1], x[5], x[9], x[13])
        ORC_QR(x[2], x[6], x[10], x[14])
        ORC_QR(x[3], x[7], x[11], x[15])
        ORC_QR(x[0], x[5], x[10], x[15])
        ORC_QR(x[1], x[6], x[11], x[12])
        ORC_QR(x[2], x[7], x[8], x[13])
        ORC_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
}

void orc_prg_fill(const uint32_t key[8], uint64_t stream, uint64_t word_offset, uint64_t* out, size_t n_words) {
    if (n_words == 0) return;
    uint64_t first_blk = word_offset / 8, last_blk = (word_offset + n_words - 1) / 8;
#pragma omp parallel for schedule(static)
    for (uint64_t b = first_blk; b <= last_blk; ++b) {
        uint32_t nonce[3] = {(uint32_t)stream, (uint32_t)(stream >> 32), (uint32_t)(b >> 32)};
        uint32_t ks[16];
        orc_chacha20_block(key, (uint32_t)b, nonce, ks);
        for (int j = 0; j < 8; ++j) {
            uint64_t w = b * 8 + (uint64_t)j;
            if (w < word_offset || w >= word_offset + n_words) continue;
            out[w - word_offset] = (uint64_t)ks[2 * j] | ((uint64_t)ks[2 * j + 1] << 32); /* little endian */
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* (3) fixed point / elementwise                                                                      */
/* ------------------------------------------------------------------------------------------------ */
uint64_t orc_encode_fixed(double x, int f) {
    /* gcn.h:191,676: static_cast<uint64_t>(x * (1<<SCALER_BIT_LENGTH)); signed values go through int64 */
    return (uint64_t)(int64_t)(x * (double)(1ull << f));
}
double orc_decode_fixed(uint64_t v, int f) { return (double)(int64_t)v / (double)(1ull << f); }

void orc_encode(const double* x, uint64_t* out, size_t n, int f) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = orc_encode_fixed(x[i], f);
}
void orc_decode(const uint64_t* v, double* out, size_t n, int f) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = orc_decode_fixed(v[i], f);
}
void orc_share_split(const double* x, size_t n, int f, const uint32_t key[8], uint64_t stream,
                     uint64_t word_offset, uint64_t* s0, uint64_t* s1) {
    orc_prg_fill(key, stream, word_offset, s1, n);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) s0[i] = orc_encode_fixed(x[i], f) - s1[i];
}
void orc_open_decode(const uint64_t* s0, const uint64_t* s1, double* out, size_t n, int f) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = orc_decode_fixed(s0[i] + s1[i], f);
}
void orc_add(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = a[i] + b[i];
}
void orc_sub(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = a[i] - b[i];
}
uint64_t orc_trunc_share(uint64_t z, int f, int share) {
    if (f <= 0) return z;
    return share == 0 ? (z >> f) : (uint64_t)0 - (((uint64_t)0 - z) >> f);
}
void orc_trunc(const uint64_t* x, uint64_t* out, size_t n, int f, int share) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = orc_trunc_share(x[i], f, share);
}
void orc_scale_public(const uint64_t* x, uint64_t c, uint64_t* out, size_t n, int f, int share) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = orc_trunc_share(x[i] * c, f, share);
}
void orc_apply_gradient(const uint64_t* W, const uint64_t* d, uint64_t lr, uint64_t* out, size_t n, int f, int share) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = W[i] - orc_trunc_share(d[i] * lr, f, share);
}
void orc_rowmul_beaver_finish(const uint64_t* e, const uint64_t* fv, const uint64_t* a, const uint64_t* b,
                              const uint64_t* c, uint64_t* out, size_t rows, size_t D, int share, int f) {
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < rows; ++r) {
        uint64_t fr = fv[r], br = b[r];
        for (size_t j = 0; j < D; ++j) {
            size_t i = r * D + j;
            uint64_t z = c[i] + e[i] * br + fr * a[i];
            if (share == 0) z += e[i] * fr;
            out[i] = orc_trunc_share(z, f, share);
        }
    }
}
void orc_cond_add(const uint64_t* v, const uint64_t* u, const uint8_t* cond, uint64_t* out, size_t rows, size_t D) {
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < rows; ++r)
        for (size_t j = 0; j < D; ++j) out[r * D + j] = v[r * D + j] + (cond[r] ? u[r * D + j] : 0);
}
void orc_transpose(const uint64_t* in, uint64_t* out, size_t rows, size_t cols) {
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < rows; ++r)
        for (size_t c = 0; c < cols; ++c) out[c * rows + r] = in[r * cols + c];
}


/* ------------------------------------------------------------------------------------------------ */
/* 2PC-RESIDUAL stand-ins (ideal functionality; see cgb_oracle.h).  Built with -ffp-contract=off.     */
/* ------------------------------------------------------------------------------------------------ */
static const double ORC_EXP_C[14] = {0x1.0000000000000p+0, 0x1.0000000000000p+0, 0x1.0000000000000p-1, 0x1.5555555555555p-3,
                                     0x1.5555555555555p-5, 0x1.1111111111111p-7, 0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-13,
                                     0x1.a01a01a01a01ap-16, 0x1.71de3a556c734p-19, 0x1.27e4fb7789f5cp-22, 0x1.ae64567f544e4p-26,
                                     0x1.1eed8eff8d898p-29, 0x1.6124613a86d09p-33};
double orc_det_exp(double x) {
    if (x < -700.0) return 0.0;
    if (x > 700.0) x = 700.0;
    const double t = x * 0x1.71547652b82fep+0 + 0.5;
    long long k = (long long)t;
    if ((double)k > t) k -= 1; /* floor */
    const double kd = (double)k;
    const double r = (x - kd * 0x1.62e42fee00000p-1) - kd * 0x1.a39ef35793c76p-33;
    double p = ORC_EXP_C[13];
    for (int i = 12; i >= 0; --i) p = p * r + ORC_EXP_C[i];
    union { uint64_t u; double d; } two_k;
    two_k.u = (uint64_t)(k + 1023) << 52; /* |k| <= 1011 */
    return p * two_k.d;
}
void orc_ideal_softmax(const uint64_t* z0, const uint64_t* z1, const int32_t* labels, size_t n, size_t C,
                       size_t train_rows, int f, uint64_t* P, uint64_t* pmy) {
    const double scale = (double)(1ull << f);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        double m = 0.0;
        for (size_t j = 0; j < C; ++j) {
            const double v = (double)(int64_t)(z0[i * C + j] + z1[i * C + j]) / scale;
            if (j == 0 || v > m) m = v;
        }
        double tot = 0.0;
        for (size_t j = 0; j < C; ++j)
            tot += orc_det_exp((double)(int64_t)(z0[i * C + j] + z1[i * C + j]) / scale - m);
        for (size_t j = 0; j < C; ++j) {
            const double e = orc_det_exp((double)(int64_t)(z0[i * C + j] + z1[i * C + j]) / scale - m);
            const uint64_t pj = (uint64_t)(int64_t)((e / tot) * scale);
            P[i * C + j] = pj;
            pmy[i * C + j] = i < train_rows ? pj - ((size_t)labels[i] == j ? (1ull << f) : 0ull) : 0ull;
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* (1) scatter / gather-sum                                                                           */
/* ------------------------------------------------------------------------------------------------ */
void orc_gather_sum_csr(const uint32_t* rowptr, const uint32_t* col, const uint64_t* x, const uint64_t* delta,
                        uint64_t* y, size_t n_rows, size_t D) {
#pragma omp parallel for schedule(dynamic, 256)
    for (size_t v = 0; v < n_rows; ++v) {
        uint64_t* yr = y + v * D;
        if (delta == y) { /* in place: y already holds delta (accumulation of the blocks of several source parties) */ }
        else if (delta) memcpy(yr, delta + v * D, D * sizeof(uint64_t));
        else memset(yr, 0, D * sizeof(uint64_t));
        for (uint32_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
            const uint64_t* xr = x + (size_t)col[e] * D;
            for (size_t j = 0; j < D; ++j) yr[j] += xr[j];
        }
    }
}
void orc_expand_rows(const uint32_t* idx, size_t n_out, const uint64_t* x, const uint64_t* delta, uint64_t* y, size_t D) {
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < n_out; ++j) {
        uint64_t* yr = y + j * D;
        if (idx[j] == ORC_NO_ROW) memset(yr, 0, D * sizeof(uint64_t));
        else memcpy(yr, x + (size_t)idx[j] * D, D * sizeof(uint64_t));
        if (delta)
            for (size_t k = 0; k < D; ++k) yr[k] += delta[j * D + k];
    }
}
void orc_segsum(const uint32_t* segptr, size_t n_seg, const uint64_t* in, uint64_t* out, size_t D, int dup) {
#pragma omp parallel for schedule(dynamic, 256)
    for (size_t s = 0; s < n_seg; ++s) {
        uint64_t acc_small[64];
        uint64_t* acc = D <= 64 ? acc_small : (uint64_t*)malloc(D * sizeof(uint64_t));
        memset(acc, 0, D * sizeof(uint64_t));
        for (uint32_t e = segptr[s]; e < segptr[s + 1]; ++e)
            for (size_t j = 0; j < D; ++j) acc[j] += in[(size_t)e * D + j];
        if (dup) {
            for (uint32_t e = segptr[s]; e < segptr[s + 1]; ++e) memcpy(out + (size_t)e * D, acc, D * sizeof(uint64_t));
        } else {
            memcpy(out + s * D, acc, D * sizeof(uint64_t));
        }
        if (acc != acc_small) free(acc);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* (2) dense contraction                                                                              */
/* ------------------------------------------------------------------------------------------------ */
void orc_matmul(const uint64_t* A, const uint64_t* B, uint64_t* C, size_t M, size_t K, size_t N, int transA, int accumulate) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < M; ++i) {
        uint64_t* c = C + i * N;
        if (!accumulate) memset(c, 0, N * sizeof(uint64_t));
        for (size_t k = 0; k < K; ++k) {
            uint64_t a = transA ? A[k * M + i] : A[i * K + k];
            const uint64_t* b = B + k * N;
            for (size_t j = 0; j < N; ++j) c[j] += a * b[j];
        }
    }
}
void orc_beaver_matmul_finish(const uint64_t* E, const uint64_t* F, const uint64_t* U, const uint64_t* V,
                              const uint64_t* Z, uint64_t* C, size_t M, size_t K, size_t N, int share, int f) {
    /* C = Z + E*(V [+ F if share 0]) + U*F */
    uint64_t* VF = (uint64_t*)malloc(K * N * sizeof(uint64_t));
    uint64_t* T = (uint64_t*)malloc(M * N * sizeof(uint64_t));
    if (share == 0) orc_add(V, F, VF, K * N);
    else memcpy(VF, V, K * N * sizeof(uint64_t));
    memcpy(T, Z, M * N * sizeof(uint64_t));
    orc_matmul(E, VF, T, M, K, N, 0, 1);
    orc_matmul(U, F, T, M, K, N, 0, 1);
    orc_trunc(T, C, M * N, f, share);
    free(VF);
    free(T);
}
