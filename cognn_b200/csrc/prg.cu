// prg.cu -- op (4): device-side PRG for correlated randomness (OM masks r/s, Beaver triples, share splitting).
//
// The reference draws its randomness inside the absent Task-Worker / SCI trees (CryptoUtil::intoShares,
// optimize-gcn/gcn.h:70,96; OM preprocessing ss_vertex_centric_algo_kernel.h:536-613).  Here it is the RFC 8439
// ChaCha20 block function used as a counter-mode PRG, so streams are reproducible bit for bit on CPU and GPU:
//   word w (u64, little endian) of stream S = bytes [8*(w%8), +8) of block b = w/8,
//   block b: key = key, counter = (u32) b, nonce = { (u32) S, (u32)(S>>32), (u32)(b>>32) }.
// One thread computes one 64-byte block in registers (ALU-bound, ~1000 integer ops per block); a warp stages its
// 32 blocks through shared memory so global stores (and the loads of the fused mask-subtract) are coalesced.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int c) { return __funnelshift_l(v, v, c); }

#define CGB_QR(a, b, c, d)              \
    a += b; d ^= a; d = rotl32(d, 16);  \
    c += d; b ^= c; b = rotl32(b, 12);  \
    a += b; d ^= a; d = rotl32(d, 8);   \
    c += d; b ^= c; b = rotl32(b, 7);

struct Key { uint32_t k[8]; };

constexpr int PRG_THREADS = 128;  // 4 warps, 32 blocks (2 KB) each

// MODE 0: out = ks;  MODE 1: out = in - ks;
// MODE 2 (2PC-RESIDUAL stand-in + re-share in one pass): out = gate(in + in2; z0 + z1) - ks, gate(v; g) = (int64) g > 0 ? v : 0,
//         z0 == nullptr: g = v (ReLU), else ReLU' of the pre-activation z applied to the gradient v
struct PrgExtra {
    const u64 *in2, *z0, *z1;
};
template <int MODE>
__global__ void __launch_bounds__(PRG_THREADS) prg_kernel(const Key key, uint64_t stream, const u64* __restrict__ stream_bias,
                                                         uint64_t word_offset, const u64* in, u64* out, uint64_t n_words,
                                                         uint64_t first_blk, uint64_t n_blks, const PrgExtra ex = PrgExtra{}) {
    if (stream_bias) stream += __ldg(stream_bias);
    __shared__ uint32_t stage[PRG_THREADS / 32][32][17];  // +1 word padding: conflict-free column reads
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t)gridDim.x * (PRG_THREADS / 32);
    for (uint64_t wb = (uint64_t)blockIdx.x * (PRG_THREADS / 32) + warp; wb * 32 < n_blks; wb += warps_total) {
        const uint64_t blk = first_blk + wb * 32 + lane;
        uint32_t s[16], x[16];
        s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
#pragma unroll
        for (int i = 0; i < 8; ++i) s[4 + i] = key.k[i];
        s[12] = (uint32_t)blk;
        s[13] = (uint32_t)stream; s[14] = (uint32_t)(stream >> 32); s[15] = (uint32_t)(blk >> 32);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = s[i];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            CGB_QR(x[0], x[4], x[8], x[12])
            CGB_QR(x[1], x[5], x[9], x[13])
            CGB_QR(x[2], x[6], x[10], x[14])
            CGB_QR(x[3], x[7], x[11], x[15])
            CGB_QR(x[0], x[5], x[10], x[15])
            CGB_QR(x[1], x[6], x[11], x[12])
            CGB_QR(x[2], x[7], x[8], x[13])
            CGB_QR(x[3], x[4], x[9], x[14])
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) stage[warp][lane][i] = x[i] + s[i];
        __syncwarp();
        // the warp's 32 blocks = 256 consecutive u64 words starting at word (first_blk + wb*32) * 8
        const uint64_t w0 = (first_blk + wb * 32) * 8;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int j = it * 32 + lane;  // word within the warp's span
            const uint64_t w = w0 + j;
            if (w >= word_offset && w < word_offset + n_words) {
                const int b = j >> 3, q = j & 7;
                const u64 ks = (u64)stage[warp][b][2 * q] | ((u64)stage[warp][b][2 * q + 1] << 32);
                const uint64_t o = w - word_offset;
                if constexpr (MODE == 2) {
                    const u64 v = in[o] + ex.in2[o];
                    const u64 gate = ex.z0 ? ex.z0[o] + ex.z1[o] : v;
                    out[o] = ((int64_t)gate > 0 ? v : 0ull) - ks;
                } else {
                    out[o] = MODE == 0 ? ks : in[o] - ks;
                }
            }
        }
        __syncwarp();
    }
}

// out = sum of n_in share vectors + sum of n_streams keystreams (word_offset 0): the GatherComp additions of one destination
// party (optimize-gcn/gcn.h:456-463) with the OM mask shares s_{p->t} regenerated in registers instead of being written to
// HBM by one launch each and read back by another.  One thread = one 64-byte block position of EVERY stream.
struct PrgSumArgs {
    const u64* in[16];
    uint64_t stream[16];
    int n_in, n_streams;
};
__device__ __forceinline__ void chacha_block(const Key& key, uint64_t stream, uint64_t blk, uint32_t* out16) {
    uint32_t s[16], x[16];
    s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[4 + i] = key.k[i];
    s[12] = (uint32_t)blk;
    s[13] = (uint32_t)stream; s[14] = (uint32_t)(stream >> 32); s[15] = (uint32_t)(blk >> 32);
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = s[i];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        CGB_QR(x[0], x[4], x[8], x[12])
        CGB_QR(x[1], x[5], x[9], x[13])
        CGB_QR(x[2], x[6], x[10], x[14])
        CGB_QR(x[3], x[7], x[11], x[15])
        CGB_QR(x[0], x[5], x[10], x[15])
        CGB_QR(x[1], x[6], x[11], x[12])
        CGB_QR(x[2], x[7], x[8], x[13])
        CGB_QR(x[3], x[4], x[9], x[14])
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) out16[i] = x[i] + s[i];
}
__global__ void __launch_bounds__(PRG_THREADS) prg_sum_kernel(const Key key, const PrgSumArgs a, const u64* __restrict__ stream_bias,
                                                             u64* out, uint64_t n_words, uint64_t n_blks) {
    const uint64_t bias = stream_bias ? __ldg(stream_bias) : 0ull;
    __shared__ u64 stage[PRG_THREADS / 32][32][9];  // +1 word padding
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t)gridDim.x * (PRG_THREADS / 32);
    for (uint64_t wb = (uint64_t)blockIdx.x * (PRG_THREADS / 32) + warp; wb * 32 < n_blks; wb += warps_total) {
        const uint64_t blk = wb * 32 + lane;
        u64 acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0;
        for (int k = 0; k < a.n_streams; ++k) {
            uint32_t w[16];
            chacha_block(key, a.stream[k] + bias, blk, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += (u64)w[2 * i] | ((u64)w[2 * i + 1] << 32);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) stage[warp][lane][i] = acc[i];
        __syncwarp();
        const uint64_t w0 = wb * 32 * 8;
        u64 v[8];  // all loads of the warp's span first, then the stores (out may alias an input at the same index only)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int j = it * 32 + lane;
            v[it] = stage[warp][j >> 3][j & 7];
        }
        for (int q = 0; q < a.n_in; ++q) {
            const u64* __restrict__ src = a.in[q];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const uint64_t w = w0 + it * 32 + lane;
                if (w < n_words) v[it] += __ldcg(src + w);
            }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const uint64_t w = w0 + it * 32 + lane;
            if (w < n_words) out[w] = v[it];
        }
        __syncwarp();
    }
}

// Several keystream fills in ONE launch (blockIdx.y = segment): everything the dealer emulation hands a side for one Beaver
// triple or one OM correlation.  Per segment: out = PRG(a) [+ PRG(b)], and optionally out_b = PRG(b) alone -- the helper's own
// share of a triple operand next to the opened sum the dealer needs for Z1 = (U0 + U1)(V0 + V1) - Z0.
struct PrgSeg {
    u64* out;
    u64* out_b;
    uint64_t n_words;
    uint64_t stream_a, stream_b;
    uint32_t has_b;
};
struct PrgMultiArgs {
    PrgSeg seg[16];
};
__global__ void __launch_bounds__(PRG_THREADS) prg_multi_kernel(const Key key, const __grid_constant__ PrgMultiArgs a,
                                                               const u64* __restrict__ stream_bias) {
    const PrgSeg& sg = a.seg[blockIdx.y];
    const uint64_t bias = stream_bias ? __ldg(stream_bias) : 0ull;
    const uint64_t n_blks = (sg.n_words + 7) / 8;
    __shared__ u64 stage[2][PRG_THREADS / 32][32][9];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t)gridDim.x * (PRG_THREADS / 32);
    for (uint64_t wb = (uint64_t)blockIdx.x * (PRG_THREADS / 32) + warp; wb * 32 < n_blks; wb += warps_total) {
        const uint64_t blk = wb * 32 + lane;
        uint32_t w[16];
        chacha_block(key, sg.stream_a + bias, blk, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) stage[0][warp][lane][i] = (u64)w[2 * i] | ((u64)w[2 * i + 1] << 32);
        if (sg.has_b) {
            chacha_block(key, sg.stream_b + bias, blk, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) stage[1][warp][lane][i] = (u64)w[2 * i] | ((u64)w[2 * i + 1] << 32);
        }
        __syncwarp();
        const uint64_t w0 = wb * 32 * 8;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int j = it * 32 + lane;
            const uint64_t wi = w0 + j;
            if (wi < sg.n_words) {
                u64 v = stage[0][warp][j >> 3][j & 7];
                if (sg.has_b) {
                    const u64 vb = stage[1][warp][j >> 3][j & 7];
                    if (sg.out_b) sg.out_b[wi] = vb;
                    v += vb;
                }
                sg.out[wi] = v;
            }
        }
        __syncwarp();
    }
}

int launch_prg(cgb_ctx* ctx, int mode, const uint32_t key[8], uint64_t stream, uint64_t word_offset, const u64* in,
               u64* out, uint64_t n_words, PrgExtra ex = PrgExtra{}) {
    if (n_words == 0) return CGB_OK;
    Key k;
    for (int i = 0; i < 8; ++i) k.k[i] = key[i];
    const uint64_t first_blk = word_offset / 8, last_blk = (word_offset + n_words - 1) / 8;
    const uint64_t n_blks = last_blk - first_blk + 1;
    uint64_t warps = (n_blks + 31) / 32;
    uint64_t blocks = (warps + (PRG_THREADS / 32) - 1) / (PRG_THREADS / 32);
    const uint64_t cap = (uint64_t)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    if (mode == 2)
        prg_kernel<2><<<(unsigned)blocks, PRG_THREADS, 0, ctx->stream>>>(k, stream, (const u64*)ctx->prg_bias, word_offset, in,
                                                                        out, n_words, first_blk, n_blks, ex);
    else if (mode == 0)
        prg_kernel<0><<<(unsigned)blocks, PRG_THREADS, 0, ctx->stream>>>(k, stream, (const u64*)ctx->prg_bias, word_offset, in,
                                                                        out, n_words, first_blk, n_blks);
    else
        prg_kernel<1><<<(unsigned)blocks, PRG_THREADS, 0, ctx->stream>>>(k, stream, (const u64*)ctx->prg_bias, word_offset, in,
                                                                        out, n_words, first_blk, n_blks);
    CGB_CHECK_LAUNCH(ctx, "prg_kernel");
    return CGB_OK;
}

}  // namespace

extern "C" {

int cgb_prg_fill(cgb_ctx* ctx, const uint32_t key[8], uint64_t stream, uint64_t word_offset, uint64_t* d_out,
                 uint64_t n_words) {
    CGB_REQUIRE(ctx, key && (d_out || n_words == 0), "cgb_prg_fill: null argument");
    return launch_prg(ctx, 0, key, stream, word_offset, nullptr, (u64*)d_out, n_words);
}
int cgb_prg_mask_sub(cgb_ctx* ctx, const uint32_t key[8], uint64_t stream, uint64_t word_offset,
                     const uint64_t* d_in, uint64_t* d_out, uint64_t n_words) {
    CGB_REQUIRE(ctx, key && ((d_in && d_out) || n_words == 0), "cgb_prg_mask_sub: null argument");
    return launch_prg(ctx, 1, key, stream, word_offset, (const u64*)d_in, (u64*)d_out, n_words);
}

int cgb_prg_fill_multi(cgb_ctx* ctx, const uint32_t key[8], const cgb_prg_seg* segs, uint32_t n_seg) {
    CGB_REQUIRE(ctx, key && (segs || n_seg == 0), "cgb_prg_fill_multi: null argument");
    for (uint32_t base = 0; base < n_seg; base += 16) {
        PrgMultiArgs a;
        memset(&a, 0, sizeof(a));
        uint32_t cnt = 0;
        uint64_t longest = 0;
        for (uint32_t j = base; j < n_seg && cnt < 16; ++j) {
            if (segs[j].n_words == 0) continue;
            CGB_REQUIRE(ctx, segs[j].out, "cgb_prg_fill_multi: null output");
            CGB_REQUIRE(ctx, segs[j].has_b || !segs[j].out_b, "cgb_prg_fill_multi: out_b without a second stream");
            a.seg[cnt].out = (u64*)segs[j].out;
            a.seg[cnt].out_b = (u64*)segs[j].out_b;
            a.seg[cnt].n_words = segs[j].n_words;
            a.seg[cnt].stream_a = segs[j].stream_a;
            a.seg[cnt].stream_b = segs[j].stream_b;
            a.seg[cnt].has_b = segs[j].has_b ? 1u : 0u;
            longest = std::max<uint64_t>(longest, segs[j].n_words);
            ++cnt;
        }
        if (cnt == 0) continue;
        Key k;
        for (int i = 0; i < 8; ++i) k.k[i] = key[i];
        const uint64_t warps = ((longest + 7) / 8 + 31) / 32;
        uint64_t bx = (warps + (PRG_THREADS / 32) - 1) / (PRG_THREADS / 32);
        const uint64_t cap = std::max<uint64_t>(1, (uint64_t)ctx->num_sms * 16 / cnt);
        if (bx > cap) bx = cap;
        prg_multi_kernel<<<dim3((unsigned)bx, cnt), PRG_THREADS, 0, ctx->stream>>>(k, a, (const u64*)ctx->prg_bias);
        CGB_CHECK_LAUNCH(ctx, "prg_multi_kernel");
    }
    return CGB_OK;
}
int cgb_ideal_relu_reshare(cgb_ctx* ctx, const uint32_t key[8], uint64_t stream, const uint64_t* d_a0, const uint64_t* d_a1,
                           const uint64_t* d_z0, const uint64_t* d_z1, uint64_t* d_out, uint64_t n_words) {
    CGB_REQUIRE(ctx, key && ((d_a0 && d_a1 && d_out) || n_words == 0), "cgb_ideal_relu_reshare: null argument");
    CGB_REQUIRE(ctx, (d_z0 == nullptr) == (d_z1 == nullptr), "cgb_ideal_relu_reshare: z0 and z1 go together");
    PrgExtra ex{(const u64*)d_a1, (const u64*)d_z0, (const u64*)d_z1};
    return launch_prg(ctx, 2, key, stream, 0, (const u64*)d_a0, (u64*)d_out, n_words, ex);
}
int cgb_prg_sum(cgb_ctx* ctx, const uint32_t key[8], const uint64_t* streams, uint32_t n_streams, const uint64_t* const* d_in,
                uint32_t n_in, uint64_t* d_out, uint64_t n_words) {
    CGB_REQUIRE(ctx, key && (d_out || n_words == 0), "cgb_prg_sum: null argument");
    CGB_REQUIRE(ctx, n_streams <= 16 && n_in <= 16 && (n_streams == 0 || streams) && (n_in == 0 || d_in),
                "cgb_prg_sum: at most 16 streams and 16 inputs");
    if (n_words == 0) return CGB_OK;
    PrgSumArgs a;
    a.n_in = (int)n_in;
    a.n_streams = (int)n_streams;
    for (uint32_t j = 0; j < n_in; ++j) {
        CGB_REQUIRE(ctx, d_in[j], "cgb_prg_sum: null input");
        a.in[j] = (const u64*)d_in[j];
    }
    for (uint32_t k = 0; k < n_streams; ++k) a.stream[k] = streams[k];
    Key k;
    for (int i = 0; i < 8; ++i) k.k[i] = key[i];
    const uint64_t n_blks = (n_words + 7) / 8;
    const uint64_t warps = (n_blks + 31) / 32;
    uint64_t blocks = (warps + (PRG_THREADS / 32) - 1) / (PRG_THREADS / 32);
    const uint64_t cap = (uint64_t)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    prg_sum_kernel<<<(unsigned)blocks, PRG_THREADS, 0, ctx->stream>>>(k, a, (const u64*)ctx->prg_bias, (u64*)d_out, n_words, n_blks);
    CGB_CHECK_LAUNCH(ctx, "prg_sum_kernel");
    return CGB_OK;
}

}  // extern "C"
