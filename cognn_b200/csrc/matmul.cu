// matmul.cu -- op (2): dense contraction mod 2^64 and the Beaver-triple recombination of the Apply step.
//
// Serves sci::twoPartyGCNMatMul (optimize-gcn/gcn.h:233, 665, 671, 710): X (N_p x F) * W (F x H), (p-y) * W^T,
// h_t * v.  This is the 64-bit integer IMAD path of the north star: every u64 multiply-add is one
// IMAD.WIDE.U32 (lo*lo) plus two 32-bit IMADs for the cross terms that land in the low 64 bits; operands are
// staged through shared memory in BK-deep tiles, each thread keeps a TM x TN block of u64 accumulators.
// The kernel takes up to two (A, B) operand pairs that accumulate into the same tile, an optional addend Z and a
// fused truncation, so the Beaver finish  C_i = trunc(Z_i + E*(V_i [+F]) + U_i*F)  is one launch.
// Small output tiles with a long K (the weight-gradient h_t * v, K = N_p) are split along K and combined with
// 64-bit atomics (addition mod 2^64 is order-independent, so this stays bit exact).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

// tensor-core (tcgen05 kind::i8 limb) path, matmul_tc.cu
int cgb_matmul_tc_run(cgb_ctx* ctx, const u64* const A[2], const u64* const B[2], int n_pairs, const u64* Z, u64* C, uint32_t M,
                      uint32_t K, uint32_t N, int transA, int accumulate, int f, int share);
#ifndef CGB_MATMUL_AUTO_TC
#define CGB_MATMUL_AUTO_TC 1  // the tensor-core path is the measured winner for wide outputs (see run_matmul)
#endif

namespace {

struct MatmulArgs {
    const u64* A[2];
    const u64* B[2];
    int n_pairs;
    const u64* Z;  // optional addend (M x N), only with k_splits == 1
    u64* C;
    uint32_t M, K, N;
    int transA;      // A stored K x M
    int accumulate;  // C += (k_splits == 1: read-modify-write; else atomics onto existing C)
    int f, share;    // fused truncation (f <= 0: none), only with k_splits == 1
    uint32_t k_chunk;  // K range per split (multiple of BK)
    uint32_t k_splits;
    size_t scratch_off;  // host side: bytes at the start of ctx->scratch the caller keeps for itself (V + F of the Beaver finish)
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), (TM * TN >= 32) ? 1 : 2) matmul_kernel(const MatmulArgs a) {
    constexpr int THREADS = (BM / TM) * (BN / TN);
    constexpr int A_LD = BM * BK / THREADS;
    constexpr int B_LD = (BK * BN + THREADS - 1) / THREADS;
    constexpr int AS = BM + 2, BS = BN + 2;  // even padding keeps 16-byte alignment of the fragment reads
    __shared__ __align__(16) u64 As[BK][AS];
    __shared__ __align__(16) u64 Bs[BK][BS];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const uint32_t m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const uint32_t kb = blockIdx.z * a.k_chunk;
    const uint32_t ke = min(a.K, kb + a.k_chunk);

    u64 acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0;

    u64 ra[A_LD], rb[B_LD];

    auto load_tiles = [&](const u64* __restrict__ A, const u64* __restrict__ B, uint32_t k0) {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            const int idx = tid + i * THREADS;
            uint32_t m, k;
            if (a.transA) { m = idx % BM; k = idx / BM; }  // consecutive threads walk M (contiguous in K x M storage)
            else { k = idx % BK; m = idx / BK; }           // consecutive threads walk K (contiguous in M x K storage)
            const uint32_t gm = m0 + m, gk = k0 + k;
            u64 v = 0;
            if (gm < a.M && gk < ke) v = a.transA ? __ldg(A + (size_t)gk * a.M + gm) : __ldg(A + (size_t)gm * a.K + gk);
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            const int idx = tid + i * THREADS;
            const uint32_t n = idx % BN, k = idx / BN;
            const uint32_t gn = n0 + n, gk = k0 + k;
            u64 v = 0;
            if (idx < BK * BN && gn < a.N && gk < ke) v = __ldg(B + (size_t)gk * a.N + gn);
            rb[i] = v;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            const int idx = tid + i * THREADS;
            int m, k;
            if (a.transA) { m = idx % BM; k = idx / BM; }
            else { k = idx % BK; m = idx / BK; }
            As[k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            const int idx = tid + i * THREADS;
            if (idx < BK * BN) Bs[idx / BN][idx % BN] = rb[i];
        }
    };

#pragma unroll
    for (int p = 0; p < 2; ++p) {
        if (p >= a.n_pairs || kb >= ke) break;
        const u64* A = p == 0 ? a.A[0] : a.A[1];
        const u64* B = p == 0 ? a.B[0] : a.B[1];
        load_tiles(A, B, kb);
        for (uint32_t k0 = kb; k0 < ke; k0 += BK) {
            __syncthreads();  // previous tile fully consumed
            store_tiles();
            __syncthreads();
            if (k0 + BK < ke) load_tiles(A, B, k0 + BK);  // prefetch next tile into registers
#pragma unroll 4
            for (int k = 0; k < BK; ++k) {
                u64 fa[TM], fb[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) fa[i] = As[k][ty * TM + i];
#pragma unroll
                for (int j = 0; j < TN; ++j) fb[j] = Bs[k][tx * TN + j];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] += fa[i] * fb[j];
            }
        }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const uint32_t gm = m0 + ty * TM + i;
        if (gm >= a.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const uint32_t gn = n0 + tx * TN + j;
            if (gn >= a.N) continue;
            const size_t o = (size_t)gm * a.N + gn;
            if (a.k_splits > 1) {
                a.C[(size_t)blockIdx.z * a.M * a.N + o] = acc[i][j];  // this split's plane; add_trunc_kernel sums the planes
            } else {
                u64 v = acc[i][j];
                if (a.Z) v += a.Z[o];
                if (a.accumulate) v += a.C[o];
                a.C[o] = trunc_share(v, a.f, a.share);
            }
        }
    }
}

// out = trunc( sum of the `splits` partial planes of t + z + (acc ? out : 0) ).  A CTA is 32 outputs x 8 plane lanes: with a
// few hundred planes over a few hundred outputs (weight gradients) the planes are what has to be spread over threads.
__global__ void __launch_bounds__(256) add_trunc_kernel(const u64* __restrict__ t, uint32_t splits, const u64* __restrict__ z,
                                                        u64* out, uint64_t n, int accumulate, int f, int share) {
    __shared__ u64 part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (uint64_t i0 = (uint64_t)blockIdx.x * 32; i0 < n; i0 += (uint64_t)gridDim.x * 32) {
        const uint64_t i = i0 + tx;
        u64 v = 0;
        if (i < n)
            for (uint32_t sp = ty; sp < splits; sp += 8) v += t[(size_t)sp * n + i];
        part[ty][tx] = v;
        __syncthreads();
        if (ty == 0 && i < n) {
#pragma unroll
            for (int k = 1; k < 8; ++k) v += part[k][tx];
            if (z) v += z[i];
            if (accumulate) v += out[i];
            out[i] = trunc_share(v, f, share);
        }
        __syncthreads();
    }
}

template <int BM, int BN, int BK, int TM, int TN>
void launch_cfg(cgb_ctx* ctx, MatmulArgs& a) {
    dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, a.k_splits);
    matmul_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, ctx->stream>>>(a);
}

// Tile shape and split-K plan of the integer-pipe kernel for one problem size.  K is split when the output tiles alone cannot
// fill the machine (weight gradients h^T v: a handful of tiles, K = vertices of the party); every split then owns a plane of
// partial sums, so the scratch need grows with the split count and is capped.
struct SplitPlan {
    uint32_t BM, BN, splits, k_chunk;
    size_t plane_bytes;  // splits * M * N * 8 when splits > 1, else 0
};
SplitPlan plan_splits(const cgb_ctx* ctx, uint32_t M, uint32_t K, uint32_t N) {
    constexpr uint32_t BK = 16;
    static const int depth_env = getenv("CGB_MATMUL_SPLIT_DEPTH") ? atoi(getenv("CGB_MATMUL_SPLIT_DEPTH")) : 0;
    static const int want_env = getenv("CGB_MATMUL_SPLIT_WANT") ? atoi(getenv("CGB_MATMUL_SPLIT_WANT")) : 0;
    static const int small_env = getenv("CGB_MATMUL_SMALL_TILES") ? atoi(getenv("CGB_MATMUL_SMALL_TILES")) : 1;
    const uint64_t depth = depth_env >= 1 && depth_env <= 64 ? (uint64_t)depth_env : 2;  // at least this many K steps per split
    const uint64_t want = (want_env > 0 ? (uint64_t)want_env : 2ull) * ctx->num_sms;
    const uint32_t sms = (uint32_t)ctx->num_sms;
    SplitPlan p{};
    // Tile shape.  Every rule below is a measurement (tools/small_kernel_probe.py, profiles/r3*_probe*.jsonl):
    if (N > 32) { p.BM = 128; p.BN = 64; }
    else if (N > 8) { p.BM = 128; p.BN = 16; }
    else { p.BM = 256; p.BN = 8; }
    if (small_env) {
        const uint32_t row_tiles = (M + 127) / 128;
        // a 64-wide tile over a few hundred row tiles pads N and quantises into waves: 21168 x 16 x 40 is 166 CTAs of 128 x 64
        // on 148 SMs (40 us); as 498 CTAs of 128 x 16 it takes 22 us.  Only when 16-wide tiles pad N clearly less.
        if (N > 32 && (uint64_t)row_tiles * ((N + 63) / 64) < 2ull * sms && ((N + 15) / 16) * 16 * 5 <= ((N + 63) / 64) * 64 * 4) p.BN = 16;
        // narrow outputs over few rows (a party of a small graph): 64-row tiles give twice the CTAs and shorter chains --
        // Cora-shaped X W0 41.7 -> 38.1 us, H W1 9.6 -> 4.5 us, g W1^T 8.4 -> 5.7 us
        if (N <= 32 && row_tiles < sms) p.BM = 64;
    }
    if (M <= 32) {  // weight gradients h^T v (M = hidden width or classes): a 128-row tile would be 3/4 padding
        p.BM = 32;
        p.BN = N > 32 ? 64 : (N > 8 ? 16 : 8);
    }
    // K is split when the output tiles alone cannot fill the machine; every split owns a plane of partial sums (<= 64 MB)
    const uint64_t tiles = (uint64_t)((M + p.BM - 1) / p.BM) * ((N + p.BN - 1) / p.BN);
    uint32_t splits = 1;
    if (tiles < want && K >= 2 * depth * BK) {
        const uint64_t s = (want + tiles - 1) / tiles;
        const uint64_t max_s = K / (depth * BK);
        const uint64_t plane = (uint64_t)M * N * sizeof(u64);
        const uint64_t max_mem = std::max<uint64_t>(1, (64ull << 20) / std::max<uint64_t>(plane, 1));
        splits = (uint32_t)std::max<uint64_t>(1, std::min(std::min(s, max_s), max_mem));
    }
    uint32_t k_chunk = (K + splits - 1) / splits;
    k_chunk = (k_chunk + BK - 1) / BK * BK;
    if (k_chunk == 0) k_chunk = BK;
    p.splits = K == 0 ? 1 : (K + k_chunk - 1) / k_chunk;
    p.k_chunk = k_chunk;
    p.plane_bytes = p.splits > 1 ? (size_t)p.splits * M * N * sizeof(u64) : 0;
    return p;
}

// Runs the (up to two pair) product with optional Z / truncation epilogue.
int run_matmul(cgb_ctx* ctx, MatmulArgs a) {
    if (a.M == 0 || a.N == 0) return CGB_OK;
    // pipe selection (DESIGN.md "Matmul design"): CGB_MATMUL_IMPL = imad | tc | auto
    {
        const char* impl = getenv("CGB_MATMUL_IMPL");
        if (ctx->matmul_impl == 1) impl = "imad";  // cgb_ctx_set_matmul_impl overrides the environment
        else if (ctx->matmul_impl == 2) impl = "tc";
        else if (ctx->matmul_impl == 0) impl = nullptr;
        bool tc = false;
        if (impl && !strcmp(impl, "tc")) tc = a.K >= 1;
        else if (impl && !strcmp(impl, "imad")) tc = false;
        else {
            // auto (measured, profiles/r1_sweep_matmul_{tc,imad}.jsonl): the tensor pipe wins 3.4-7.4x once the 64-wide
            // tile is mostly used and there are enough 128 x 64 tiles to fill the machine; small-N outputs (H = 16, C = 7)
            // and long-K / few-tile weight gradients stay on the integer pipe (split-K)
            const uint64_t tiles = (uint64_t)((a.M + 127) / 128) * ((a.N + 63) / 64);
            tc = CGB_MATMUL_AUTO_TC && a.N >= 48 && a.K >= 32 && tiles >= (uint64_t)ctx->num_sms / 2;
        }
        if (tc) return cgb_matmul_tc_run(ctx, a.A, a.B, a.n_pairs, a.Z, a.C, a.M, a.K, a.N, a.transA, a.accumulate, a.f, a.share);
    }
    SplitPlan pl = plan_splits(ctx, a.M, a.K, a.N);
    a.k_chunk = pl.k_chunk;
    a.k_splits = pl.splits;

    const u64* Z = a.Z;
    u64* C = a.C;
    const int f = a.f, share = a.share, accumulate = a.accumulate;
    if (pl.splits > 1) {
        // every split stores its own M x N plane (plain stores: no zeroing pass, no atomics on shared addresses); the
        // finishing launch sums the planes, adds Z / the old C and truncates
        int rc = cgb_scratch_reserve(ctx, a.scratch_off + pl.plane_bytes);
        if (rc) return rc;
        a.C = (u64*)((char*)ctx->scratch + a.scratch_off);
        a.Z = nullptr; a.f = 0; a.accumulate = 0;
    }
    if (pl.BM == 32) {
        if (pl.BN == 64) launch_cfg<32, 64, 16, 2, 4>(ctx, a);
        else if (pl.BN == 16) launch_cfg<32, 16, 16, 2, 1>(ctx, a);
        else launch_cfg<32, 8, 16, 1, 1>(ctx, a);
    } else if (pl.BM == 64) {
        if (pl.BN == 16) launch_cfg<64, 16, 16, 2, 2>(ctx, a);
        else launch_cfg<64, 8, 16, 2, 1>(ctx, a);
    } else if (pl.BN == 64) launch_cfg<128, 64, 16, 8, 4>(ctx, a);
    else if (pl.BN == 16) launch_cfg<128, 16, 16, 4, 2>(ctx, a);  // (128x16 tm8 / tn4 and 256x16 tiles measured slower: r2v, r3e probes)
    else launch_cfg<256, 8, 16, 8, 1>(ctx, a);
    CGB_CHECK_LAUNCH(ctx, "matmul_kernel");
    ctx->last_kernel = pl.BM == 32 ? "matmul_kernel<32,*,16,*,*> (IMAD.WIDE u64 tiles, short M)"
                     : pl.BM == 64 ? "matmul_kernel<64,*,16,2,*> (IMAD.WIDE u64 tiles, few rows)" : pl.BN == 64 ? "matmul_kernel<128,64,16,8,4> (IMAD.WIDE u64 tiles)"
                                   : (pl.BN == 16 ? "matmul_kernel<128,16,16,4,2> (IMAD.WIDE u64 tiles)"
                                                  : "matmul_kernel<256,8,16,8,1> (IMAD.WIDE u64 tiles)");
    if (pl.splits > 1) {
        const uint64_t n = (uint64_t)a.M * a.N;
        unsigned blocks = (unsigned)std::min<uint64_t>((n + 31) / 32, (uint64_t)ctx->num_sms * 16);
        add_trunc_kernel<<<blocks, 256, 0, ctx->stream>>>((const u64*)a.C, pl.splits, Z, C, n, accumulate, f, share);
        CGB_CHECK_LAUNCH(ctx, "add_trunc_kernel");
    }
    return CGB_OK;
}

}  // namespace

extern "C" {

int cgb_matmul(cgb_ctx* ctx, const uint64_t* d_A, const uint64_t* d_B, uint64_t* d_C, uint32_t M, uint32_t K,
               uint32_t N, int transA, int accumulate) {
    CGB_REQUIRE(ctx, (d_A && d_B && d_C) || M == 0 || N == 0 || K == 0, "cgb_matmul: null argument");
    CGB_REQUIRE(ctx, (const void*)d_C != (const void*)d_A && (const void*)d_C != (const void*)d_B,
                "cgb_matmul: C must not alias A or B");
    MatmulArgs a{};
    a.A[0] = (const u64*)d_A; a.B[0] = (const u64*)d_B; a.n_pairs = 1;
    a.Z = nullptr; a.C = (u64*)d_C; a.M = M; a.K = K; a.N = N;
    a.transA = transA; a.accumulate = accumulate; a.f = 0; a.share = 0;
    if (K == 0) {
        if (!accumulate && M && N) CGB_CHECK_CUDA(ctx, cudaMemsetAsync(d_C, 0, (size_t)M * N * sizeof(u64), ctx->stream));
        return CGB_OK;
    }
    return run_matmul(ctx, a);
}

int cgb_beaver_matmul_finish(cgb_ctx* ctx, const uint64_t* d_E, const uint64_t* d_F, const uint64_t* d_U,
                             const uint64_t* d_V, const uint64_t* d_Z, uint64_t* d_C, uint32_t M, uint32_t K,
                             uint32_t N, int share, int f) {
    CGB_REQUIRE(ctx, d_E && d_F && d_U && d_V && d_Z && d_C, "cgb_beaver_matmul_finish: null argument");
    CGB_REQUIRE(ctx, share == 0 || share == 1, "cgb_beaver_matmul_finish: bad share");
    CGB_REQUIRE(ctx, f < 64, "cgb_beaver_matmul_finish: bad f");
    CGB_REQUIRE(ctx, d_C != d_E && d_C != d_F && d_C != d_U && d_C != d_V && d_C != d_Z,
                "cgb_beaver_matmul_finish: C must not alias the inputs");
    if (M == 0 || N == 0) return CGB_OK;
    // share 0 folds the E*F term into the first product: E*(V+F) + U*F
    const u64* Bfirst = (const u64*)d_V;
    size_t vf_bytes = 0;
    if (share == 0) {
        // keep V+F at the head of the scratch buffer, before the split-K planes (reserved together: growing the buffer later
        // would move it)
        vf_bytes = ((size_t)K * N * sizeof(u64) + 255) & ~(size_t)255;
        int rc = cgb_scratch_reserve(ctx, vf_bytes + plan_splits(ctx, M, K, N).plane_bytes);
        if (rc) return rc;
        u64* VF = (u64*)ctx->scratch;
        rc = cgb_add(ctx, d_V, d_F, (uint64_t*)VF, (uint64_t)K * N);
        if (rc) return rc;
        Bfirst = VF;
    }
    MatmulArgs a{};
    a.scratch_off = vf_bytes;
    a.A[0] = (const u64*)d_E; a.B[0] = Bfirst;
    a.A[1] = (const u64*)d_U; a.B[1] = (const u64*)d_F;
    a.n_pairs = 2;
    a.Z = (const u64*)d_Z; a.C = (u64*)d_C; a.M = M; a.K = K; a.N = N;
    a.transA = 0; a.accumulate = 0; a.f = f; a.share = share;
    return run_matmul(ctx, a);
}

int cgb_mm_open_launch(cgb_ctx* ctx, uint64_t* d_mine, const uint64_t* d_peer, uint64_t nEF, uint64_t nF, const uint64_t* d_V,
                       uint64_t* d_VF);  // elementwise.cu

int cgb_beaver_matmul_finish_open(cgb_ctx* ctx, uint64_t* d_mine, const uint64_t* d_peer, const uint64_t* d_U,
                                  const uint64_t* d_V, const uint64_t* d_Z, uint64_t* d_C, uint32_t M, uint32_t K,
                                  uint32_t N, int share, int f) {
    CGB_REQUIRE(ctx, d_mine && d_peer && d_U && d_V && d_Z && d_C, "cgb_beaver_matmul_finish_open: null argument");
    CGB_REQUIRE(ctx, share == 0 || share == 1, "cgb_beaver_matmul_finish_open: bad share");
    CGB_REQUIRE(ctx, f < 64, "cgb_beaver_matmul_finish_open: bad f");
    CGB_REQUIRE(ctx, d_C != d_mine && d_C != d_peer && d_C != d_U && d_C != d_V && d_C != d_Z,
                "cgb_beaver_matmul_finish_open: C must not alias the inputs");
    if (M == 0 || N == 0) return CGB_OK;
    const uint64_t nE = (uint64_t)M * K, nF = (uint64_t)K * N;
    u64* VF = nullptr;
    size_t vf_bytes = 0;
    if (share == 0) {  // V + F at the head of the scratch buffer, the split-K planes behind it (reserved together)
        vf_bytes = ((size_t)nF * sizeof(u64) + 255) & ~(size_t)255;
        int rc = cgb_scratch_reserve(ctx, vf_bytes + plan_splits(ctx, M, K, N).plane_bytes);
        if (rc) return rc;
        VF = (u64*)ctx->scratch;
    }
    int rc = cgb_mm_open_launch(ctx, d_mine, d_peer, nE + nF, nF, d_V, (uint64_t*)VF);
    if (rc) return rc;
    MatmulArgs a{};
    a.scratch_off = vf_bytes;
    a.A[0] = (const u64*)d_mine; a.B[0] = share == 0 ? VF : (const u64*)d_V;
    a.A[1] = (const u64*)d_U; a.B[1] = (const u64*)d_mine + nE;
    a.n_pairs = 2;
    a.Z = (const u64*)d_Z; a.C = (u64*)d_C; a.M = M; a.K = K; a.N = N;
    a.transA = 0; a.accumulate = 0; a.f = f; a.share = share;
    return run_matmul(ctx, a);
}

}  // extern "C"

// ---- pipe-ceiling probe: independent u64 multiply-adds in registers, no memory traffic (the denominator of the integer-pipe
// fraction bench.py reports; MEASURED_PEAKS.json only has HBM and bf16) ------------------------------------------------------
namespace {
template <int TM, int TN>
__global__ void __launch_bounds__(256) imad_probe_kernel(u64* out, int iters, u64 seed) {
    u64 acc[TM][TN], fa[TM], fb[TN];
    const u64 t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < TM; ++i) fa[i] = seed * (t + i + 1);
#pragma unroll
    for (int j = 0; j < TN; ++j) fb[j] = (seed ^ 0x9E3779B97F4A7C15ull) * (t + 7 * j + 3);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] += fa[i] * fb[j];
#pragma unroll
        for (int i = 0; i < TM; ++i) fa[i] += (u64)it;  // operands change every iteration: nothing can be hoisted
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) s ^= acc[i][j];
    out[t] = s;
}
}  // namespace

extern "C" int cgb_probe_imad_peak(cgb_ctx* ctx, double* u64_mac_per_s) {
    CGB_REQUIRE(ctx, u64_mac_per_s, "cgb_probe_imad_peak: null argument");
    CGB_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->num_sms * 8, threads = 256, iters = 4096;
    int rc = cgb_scratch_reserve(ctx, (size_t)blocks * threads * sizeof(u64));
    if (rc) return rc;
    cudaEvent_t e0, e1;
    CGB_CHECK_CUDA(ctx, cudaEventCreate(&e0));
    CGB_CHECK_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CGB_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        imad_probe_kernel<8, 4><<<blocks, threads, 0, ctx->stream>>>((u64*)ctx->scratch, iters, 0x1234567ull + rep);
        CGB_CHECK_LAUNCH(ctx, "imad_probe_kernel");
        CGB_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        CGB_CHECK_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        CGB_CHECK_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms > 0.f) best = std::max(best, (double)blocks * threads * iters * 32.0 / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *u64_mac_per_s = best;
    return CGB_OK;
}
