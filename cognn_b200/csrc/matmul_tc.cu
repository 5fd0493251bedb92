// matmul_tc.cu -- op (2) on the 5th-generation tensor cores: u64 x u64 -> low 64 bits through int8 limbs.
//
// Serves sci::twoPartyGCNMatMul (optimize-gcn/gcn.h:233,665,671,710) for shapes where the tensor pipe beats the
// integer pipe (N >= 64).  Every u64 is cut into 8 unsigned byte limbs; C mod 2^64 only needs the 36 limb products
// A_i * B_j with i + j <= 7, and all products of one anti-diagonal d = i + j carry the same weight 2^(8d):
//     C = sum_{d=0..7} 2^(8d) * sum_{i+j=d} A_i B_j        (mod 2^64)
// Each A_i B_j is a u8 x u8 -> s32 GEMM: tcgen05.mma kind::i8 with the accumulator of diagonal d in tensor memory
// (8 accumulators x 64 columns = all 512 TMEM columns of the SM).  Only the low 64 - 8d bits of accumulator d matter,
// so 32-bit wrap-around of the high diagonals is harmless; the low diagonals stay exact for K_total <= 4096.
//
// Data path: a pre-pass writes the limb planes to global memory already in the UMMA canonical (no-swizzle, K-major)
// shared-memory layout, one contiguous 32 KB (A: 128 rows) / 16 KB (B: 64 columns) block per (tile, 32-wide k-step),
// so the main kernel moves operands with plain 1-D bulk TMA copies (cp.async.bulk, SASS UBLKCP) -- no tensor maps.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane; the 36 limb products of a k-step are issued as 12 tcgen05.mma of N = 64..256), warps 2-5 =
// epilogue (tcgen05.ld, recombination of the 8 diagonals into u64, optional + Z and fixed-point truncation, store).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int TC_BM = 128;      // rows of A per CTA tile (= UMMA M, one TMEM lane per row)
constexpr int TC_BN = 64;       // columns of B per CTA tile (= UMMA N); 8 diagonals x 64 = 512 TMEM columns
constexpr int TC_BK = 32;       // k-step: 32 bytes per limb plane row = UMMA K for 8-bit operands
constexpr int TC_STAGES = 4;
constexpr uint32_t A_STAGE_BYTES = 8 * TC_BM * TC_BK;  // 8 limb planes
constexpr uint32_t B_STAGE_BYTES = 8 * TC_BN * TC_BK;
constexpr uint32_t A_PLANE_BYTES = TC_BM * TC_BK;
constexpr uint32_t B_PLANE_BYTES = TC_BN * TC_BK;
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024;

// ---- limb planes in the UMMA canonical layout ------------------------------------------------------------------
// K-major, SWIZZLE_NONE ("interleave"): in 16-byte units the operand tile is ((8,n),2):((1,SBO),LBO): a core matrix is
// 8 rows x 16 bytes stored contiguously (128 B), row groups follow at SBO = 128 B, the second 16-byte k-chunk of the
// 32-byte k-step at LBO = rows * 16 B.  Byte (row r, k-byte kb) of a plane therefore sits at (kb/16 * rows + r) * 16 + kb%16.
// Global layout: [row block][k-step][limb plane][that plane's tile], so one (row block, k-step) is one contiguous block.
template <int ROWS>
__device__ __forceinline__ size_t plane_offset(uint32_t rb, uint32_t ks, uint32_t n_ksteps, uint32_t plane, uint32_t r, uint32_t kb) {
    return (((size_t)rb * n_ksteps + ks) * 8 + plane) * (size_t)(ROWS * TC_BK) + (size_t)(kb >> 4) * ROWS * 16 + (size_t)r * 16 + (kb & 15);
}

// A operand: rows = output rows.  src is M x K row-major (or K x M if transA); one thread handles 16 consecutive k of one row
// and writes 8 x 16 bytes (one per limb plane).  Padding rows / columns are written as zeros.
__global__ void __launch_bounds__(256) limb_split_rows_kernel(const u64* __restrict__ src, uint8_t* __restrict__ dst, uint32_t M, uint32_t K,
                                                              uint32_t lda, uint32_t Mpad, uint32_t ks0, uint32_t n_ksteps_total,
                                                              uint32_t n_ksteps_this, int transA) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t chunks_per_row = n_ksteps_this * 2;  // 16-element chunks
    if (idx >= (uint64_t)Mpad * chunks_per_row) return;
    // consecutive threads walk rows (so that 8 threads complete a 128-byte core matrix and stores coalesce)
    const uint32_t m = (uint32_t)(idx % Mpad), ch = (uint32_t)(idx / Mpad);
    const uint32_t k0 = ch * 16;
    uint32_t w[8][4];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) w[p][q] = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t k = k0 + j;
        u64 v = 0;
        if (m < M && k < K) v = transA ? __ldg(src + (size_t)k * lda + m) : __ldg(src + (size_t)m * lda + k);
#pragma unroll
        for (int p = 0; p < 8; ++p) w[p][j >> 2] |= (uint32_t)((v >> (8 * p)) & 0xFF) << (8 * (j & 3));
    }
    const uint32_t rb = m / TC_BM, r = m % TC_BM;
    const uint32_t ks = ks0 + ch / 2, kb = (ch & 1) * 16;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        uint4* out = reinterpret_cast<uint4*>(dst + plane_offset<TC_BM>(rb, ks, n_ksteps_total, p, r, kb));
        *out = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
    }
}

// B operand: src is K x N row-major; the UMMA B operand is N x K, K-major.  One thread = one column n, 16 consecutive k.
// Inside a (column block, k-step) block the order is [k-chunk][limb plane][column][16 B]: the 8 planes form one 512-row
// K-major operand, so ONE tcgen05.mma can multiply an A plane with several consecutive B planes (N up to 256) and write
// the consecutive diagonals' accumulators -- 12 MMAs per k-step instead of 36, each A plane read from smem 1-2 times
// instead of up to 8.
__global__ void __launch_bounds__(256) limb_split_cols_kernel(const u64* __restrict__ src, uint8_t* __restrict__ dst, uint32_t K, uint32_t N,
                                                              uint32_t Npad, uint32_t ks0, uint32_t n_ksteps_total, uint32_t n_ksteps_this) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t chunks = n_ksteps_this * 2;
    if (idx >= (uint64_t)Npad * chunks) return;
    const uint32_t n = (uint32_t)(idx % Npad), ch = (uint32_t)(idx / Npad);  // consecutive threads walk n: coalesced reads of B rows
    const uint32_t k0 = ch * 16;
    uint32_t w[8][4];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) w[p][q] = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t k = k0 + j;
        u64 v = 0;
        if (n < N && k < K) v = __ldg(src + (size_t)k * N + n);
#pragma unroll
        for (int p = 0; p < 8; ++p) w[p][j >> 2] |= (uint32_t)((v >> (8 * p)) & 0xFF) << (8 * (j & 3));
    }
    const uint32_t nb = n / TC_BN, r = n % TC_BN;
    const uint32_t ks = ks0 + ch / 2, kb = (ch & 1) * 16;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const size_t off = ((size_t)nb * n_ksteps_total + ks) * (size_t)(8 * TC_BN * TC_BK) + (size_t)(kb >> 4) * (8 * TC_BN * 16) +
                           ((size_t)p * TC_BN + r) * 16;
        uint4* out = reinterpret_cast<uint4*>(dst + off);
        *out = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
    }
}

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version = 1 [46,48), layout SWIZZLE_NONE
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}

struct TcArgs {
    const uint8_t* A;  // limb planes, [row block][k-step][plane][tile]
    const uint8_t* B;
    const u64* Z;
    u64* C;
    uint32_t M, N, n_ksteps;
    int f, share, accumulate;
};

__global__ void __launch_bounds__(192, 1) matmul_tc_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], acc_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t mb = blockIdx.y, nb = blockIdx.x;
    uint8_t* sA = smem;
    uint8_t* sB = smem + TC_STAGES * A_STAGE_BYTES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&acc_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {  // TMEM allocation is a warp-wide operation; the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint8_t* gA = a.A + (size_t)mb * a.n_ksteps * A_STAGE_BYTES;
            const uint8_t* gB = a.B + (size_t)nb * a.n_ksteps * B_STAGE_BYTES;
            for (uint32_t ks = 0; ks < a.n_ksteps; ++ks) {
                const uint32_t s = ks % TC_STAGES, ph = (ks / TC_STAGES) & 1;
                mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);  // slot free (first round passes immediately)
                mbar_expect_tx(smem_u32(&full_bar[s]), A_STAGE_BYTES + B_STAGE_BYTES);
                bulk_g2s(smem_u32(sA + s * A_STAGE_BYTES), gA + (size_t)ks * A_STAGE_BYTES, A_STAGE_BYTES, smem_u32(&full_bar[s]));
                bulk_g2s(smem_u32(sB + s * B_STAGE_BYTES), gB + (size_t)ks * B_STAGE_BYTES, B_STAGE_BYTES, smem_u32(&full_bar[s]));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one lane issues every tcgen05.mma of the tile =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): c = S32 (2) [4,6), a = b = UINT8 (0), K-major both, N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc0 = (2u << 4) | ((uint32_t)(TC_BM >> 4) << 24);
            for (uint32_t ks = 0; ks < a.n_ksteps; ++ks) {
                const uint32_t s = ks % TC_STAGES, ph = (ks / TC_STAGES) & 1;
                mbar_wait(smem_u32(&full_bar[s]), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(sA + s * A_STAGE_BYTES), b0 = smem_u32(sB + s * B_STAGE_BYTES);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint64_t da = smem_desc(a0 + i * A_PLANE_BYTES, TC_BM * 16, 128);
                    // A_i times the B planes j = 0 .. 7-i lands on the diagonals i .. 7: consecutive TMEM columns.  One MMA
                    // covers up to 4 planes (N = 256); plane j of B starts at row j * 64 of the 512-row operand.
#pragma unroll
                    for (int j0 = 0; j0 + i < 8; j0 += 4) {
                        const int planes = (8 - i - j0) < 4 ? (8 - i - j0) : 4;
                        const uint32_t n = (uint32_t)planes * TC_BN;
                        const uint64_t db = smem_desc(b0 + (uint32_t)j0 * TC_BN * 16, 8 * TC_BN * 16, 128);
                        // the products with A_0 touch every diagonal first: they initialise the accumulators at the first k-step
                        tc_mma_i8(tmem_base + (uint32_t)(i + j0) * TC_BN, da, db, idesc0 | ((n >> 3) << 17), (ks > 0 || i > 0) ? 1u : 0u);
                    }
                }
                tc_commit(smem_u32(&empty_bar[s]));  // frees the smem stage once these MMAs have read it
            }
            tc_commit(smem_u32(&acc_bar));  // all accumulators final
        }
    } else {
        // ===== epilogue: warps 2..5, one output row per thread (TMEM lane = 32 * (warp % 4) + lane) =====
        mbar_wait(smem_u32(&acc_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t q = warp & 3;
        const uint32_t row_in_tile = q * 32 + lane;
        const uint32_t gm = mb * TC_BM + row_in_tile;
        const uint32_t lane_addr = tmem_base + ((q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < TC_BN / 16; ++c) {
            u64 acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0;
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                uint32_t r[16];
                const uint32_t taddr = lane_addr + (uint32_t)(d * TC_BN + c * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += (u64)r[j] << (8 * d);  // 32-bit wrap of the high diagonals only touches bits >= 64
            }
            if (gm < a.M) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t gn = nb * TC_BN + c * 16 + j;
                    if (gn < a.N) {
                        const size_t o = (size_t)gm * a.N + gn;
                        u64 v = acc[j];
                        if (a.Z) v += a.Z[o];
                        if (a.accumulate) v += a.C[o];
                        a.C[o] = trunc_share(v, a.f, a.share);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- persistent form (round 2) -------------------------------------------------------------------------------------------------
// The kernel above runs one tile per CTA: TMEM allocation, barrier set-up and the fill of the operand ring are paid per tile, and
// the tensor pipe idles while the epilogue drains the 512 accumulator columns (ncu round 1: 54 % tensor-pipe active at K = 512,
// far less at K = 128).  Here one CTA per SM walks over tiles (consecutive tiles share the A row block, so its planes stay in L2):
//   * the TMA producer runs AHEAD across tile boundaries -- while tile i is in its epilogue the ring already holds the first
//     k-steps of tile i + 1;
//   * the epilogue first pulls ALL accumulators into registers (64 u64 per thread, two tcgen05.ld per wait), releases TMEM
//     (acc_empty) and only then adds Z / truncates / stores, so the MMAs of tile i + 1 overlap the output phase of tile i;
//   * the output goes through a padded shared-memory staging chunk so that a warp stores 2 x 128 contiguous bytes per
//     instruction instead of 32 scattered 16-byte pieces (and reads Z / C the same way).
constexpr uint32_t TC_EPI_STRIDE = 17;                                  // u64 per staged row (16 + 1 pad: 2-way bank conflicts at most)
constexpr uint32_t TC_EPI_BYTES = TC_BM * TC_EPI_STRIDE * 8;            // one 128 x 16 chunk
constexpr uint32_t TC_P_SMEM_BYTES = TC_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + TC_EPI_BYTES + 1024;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct TcPArgs {
    TcArgs t;
    uint32_t n_mb, n_nb;
};

__global__ void __launch_bounds__(192, 1) matmul_tc_persistent_kernel(const __grid_constant__ TcPArgs pa) {
    const TcArgs& a = pa.t;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], acc_full, acc_empty;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sA = smem;
    uint8_t* sB = smem + TC_STAGES * A_STAGE_BYTES;
    u64* sE = reinterpret_cast<u64*>(smem + TC_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
    const uint32_t n_tiles = pa.n_mb * pa.n_nb;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&acc_full), 1);
        mbar_init(smem_u32(&acc_empty), 4);  // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer: one running stage counter over all tiles of this CTA =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const uint32_t mb = tile / pa.n_nb, nb = tile - mb * pa.n_nb;
                const uint8_t* gA = a.A + (size_t)mb * a.n_ksteps * A_STAGE_BYTES;
                const uint8_t* gB = a.B + (size_t)nb * a.n_ksteps * B_STAGE_BYTES;
                for (uint32_t ks = 0; ks < a.n_ksteps; ++ks, ++it) {
                    const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                    mbar_expect_tx(smem_u32(&full_bar[s]), A_STAGE_BYTES + B_STAGE_BYTES);
                    bulk_g2s(smem_u32(sA + s * A_STAGE_BYTES), gA + (size_t)ks * A_STAGE_BYTES, A_STAGE_BYTES, smem_u32(&full_bar[s]));
                    bulk_g2s(smem_u32(sB + s * B_STAGE_BYTES), gB + (size_t)ks * B_STAGE_BYTES, B_STAGE_BYTES, smem_u32(&full_bar[s]));
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc0 = (2u << 4) | ((uint32_t)(TC_BM >> 4) << 24);
            uint32_t it = 0, j = 0;
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
                mbar_wait(smem_u32(&acc_empty), (j & 1) ^ 1);  // the epilogue has read the previous tile out of TMEM
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (uint32_t ks = 0; ks < a.n_ksteps; ++ks, ++it) {
                    const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(sA + s * A_STAGE_BYTES), b0 = smem_u32(sB + s * B_STAGE_BYTES);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint64_t da = smem_desc(a0 + i * A_PLANE_BYTES, TC_BM * 16, 128);
#pragma unroll
                        for (int j0 = 0; j0 + i < 8; j0 += 4) {
                            const int planes = (8 - i - j0) < 4 ? (8 - i - j0) : 4;
                            const uint32_t n = (uint32_t)planes * TC_BN;
                            const uint64_t db = smem_desc(b0 + (uint32_t)j0 * TC_BN * 16, 8 * TC_BN * 16, 128);
                            tc_mma_i8(tmem_base + (uint32_t)(i + j0) * TC_BN, da, db, idesc0 | ((n >> 3) << 17), (ks > 0 || i > 0) ? 1u : 0u);
                        }
                    }
                    tc_commit(smem_u32(&empty_bar[s]));
                }
                tc_commit(smem_u32(&acc_full));
            }
        }
    } else {
        // ===== epilogue: warps 2..5 =====
        const uint32_t q = warp & 3;                  // TMEM lane quarter this warp may read
        const uint32_t et = q * 32 + lane;            // 0..127: row of the tile this thread drains
        const uint32_t lane_addr = tmem_base + ((q * 32) << 16);
        uint32_t j = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            const uint32_t mb = tile / pa.n_nb, nb = tile - mb * pa.n_nb;
            mbar_wait(smem_u32(&acc_full), j & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            u64 acc[TC_BN];
#pragma unroll
            for (int c = 0; c < TC_BN; ++c) acc[c] = 0;
#pragma unroll
            for (int c = 0; c < TC_BN / 16; ++c) {
#pragma unroll
                for (int d = 0; d < 8; d += 2) {
                    uint32_t r0[16], r1[16];
                    const uint32_t t0 = lane_addr + (uint32_t)(d * TC_BN + c * 16), t1 = t0 + TC_BN;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]), "=r"(r0[8]),
                          "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15])
                        : "r"(t0));
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]),
                          "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
                        : "r"(t1));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        acc[c * 16 + k] += ((u64)r0[k] << (8 * d)) + ((u64)r1[k] << (8 * (d + 1)));
                }
            }
            // TMEM is drained: hand it back before the output phase
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty));
            // output phase, 16 columns at a time through the staging chunk
#pragma unroll
            for (int c = 0; c < TC_BN / 16; ++c) {
#pragma unroll
                for (int k = 0; k < 16; ++k) sE[et * TC_EPI_STRIDE + k] = acc[c * 16 + k];
                asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t row = (et >> 4) + 8 * i, col = et & 15;
                    const uint32_t gm = mb * TC_BM + row, gn = nb * TC_BN + c * 16 + col;
                    if (gm < a.M && gn < a.N) {
                        const size_t o = (size_t)gm * a.N + gn;
                        u64 v = sE[row * TC_EPI_STRIDE + col];
                        if (a.Z) v += a.Z[o];
                        if (a.accumulate) v += a.C[o];
                        a.C[o] = trunc_share(v, a.f, a.share);
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- pipe-ceiling probe ------------------------------------------------------------------------------------------------
// The same 12 tcgen05.mma per k-step as matmul_tc_kernel (36 limb products, M = 128, N = 64..256, K = 32), issued back to back
// on whatever the (zeroed) shared memory holds: no TMA, no epilogue.  One CTA per SM.  Its rate is the denominator of the
// tensor-pipe fraction bench.py reports for the limb matmul: what the instruction mix could reach if operands were free.
__global__ void __launch_bounds__(64, 1) tc_pipe_probe_kernel(uint32_t n_ksteps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_base_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < (A_STAGE_BYTES + B_STAGE_BYTES) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&done_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc0 = (2u << 4) | ((uint32_t)(TC_BM >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + A_STAGE_BYTES);
        for (uint32_t ks = 0; ks < n_ksteps; ++ks) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint64_t da = smem_desc(a0 + i * A_PLANE_BYTES, TC_BM * 16, 128);
#pragma unroll
                for (int j0 = 0; j0 + i < 8; j0 += 4) {
                    const int planes = (8 - i - j0) < 4 ? (8 - i - j0) : 4;
                    const uint32_t n = (uint32_t)planes * TC_BN;
                    const uint64_t db = smem_desc(b0 + (uint32_t)j0 * TC_BN * 16, 8 * TC_BN * 16, 128);
                    tc_mma_i8(tmem_base + (uint32_t)(i + j0) * TC_BN, da, db, idesc0 | ((n >> 3) << 17), (ks > 0 || i > 0) ? 1u : 0u);
                }
            }
        }
        tc_commit(smem_u32(&done_bar));
        mbar_wait(smem_u32(&done_bar), 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

#ifdef CGB_EXPERIMENTAL_TC_MC
// ---- EXPERIMENTAL, NOT PART OF THE PRODUCT BUILD (compile with -DCGB_EXPERIMENTAL_TC_MC) ------------------------------------
// Next step named in DESIGN.md section 8: the kernel above keeps the tensor pipe 54 % busy because each 128x64 tile pulls
// 48 KB of limb planes per k-step through the L2 -> SM path, which is at its cap chip-wide.  Here the four CTAs that compute
// the four N-tiles of the same 128-row block form a cluster and SHARE the A planes: CTA r loads only slice r (8 KB, two limb
// planes) of each A stage and multicasts it into the same shared-memory offset of all four CTAs, so every SM pulls 8 + 16 KB
// per k-step instead of 32 + 16 KB.  Written and compile-checked in round 1 after the GPU budget was spent: it has NOT run on
// hardware yet and is therefore not selectable in the product library.
//
// Barrier protocol per stage s (all barriers live at the same shared-memory offset in the four CTAs):
//   full_bar[s]   count 1: the CTA's own producer arrives with expect_tx = 32 KB (A, delivered by four multicasts, one from
//                 each CTA of the cluster, its own included) + 16 KB (its own B).  A peer's bytes may land before the local
//                 expect_tx of the same round; the phase still cannot complete before the local arrive.
//   empty_bar[s]  count MC: every CTA's MMA warp commits with a multicast arrive on the four CTAs' empty_bar[s], so a producer
//                 overwrites slice r of stage s everywhere only after all four CTAs have finished reading that stage.
// The cluster is synchronised after barrier initialisation (before the first remote operation) and before exit (no CTA may
// leave while peers can still arrive on its barriers).
constexpr int TC_MC = 4;
constexpr uint32_t A_SLICE_BYTES = A_STAGE_BYTES / TC_MC;

__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(TC_MC, 1, 1) __launch_bounds__(192, 1) matmul_tc_mc_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], acc_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t mb = blockIdx.y, nb = blockIdx.x;  // the cluster spans four consecutive nb of one mb
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint16_t all = (uint16_t)((1u << TC_MC) - 1);
    uint8_t* sA = smem;
    uint8_t* sB = smem + TC_STAGES * A_STAGE_BYTES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), TC_MC);
        }
        mbar_init(smem_u32(&acc_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // every CTA's barriers are initialised before any peer multicasts into it
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            const uint8_t* gA = a.A + (size_t)mb * a.n_ksteps * A_STAGE_BYTES + (size_t)rank * A_SLICE_BYTES;
            const uint8_t* gB = a.B + (size_t)nb * a.n_ksteps * B_STAGE_BYTES;
            for (uint32_t ks = 0; ks < a.n_ksteps; ++ks) {
                const uint32_t s = ks % TC_STAGES, ph = (ks / TC_STAGES) & 1;
                mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);  // all four CTAs have released stage s
                mbar_expect_tx(smem_u32(&full_bar[s]), A_STAGE_BYTES + B_STAGE_BYTES);
                bulk_g2s_multicast(smem_u32(sA + s * A_STAGE_BYTES + rank * A_SLICE_BYTES), gA + (size_t)ks * A_STAGE_BYTES, A_SLICE_BYTES,
                                   smem_u32(&full_bar[s]), all);
                bulk_g2s(smem_u32(sB + s * B_STAGE_BYTES), gB + (size_t)ks * B_STAGE_BYTES, B_STAGE_BYTES, smem_u32(&full_bar[s]));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc0 = (2u << 4) | ((uint32_t)(TC_BM >> 4) << 24);
            for (uint32_t ks = 0; ks < a.n_ksteps; ++ks) {
                const uint32_t s = ks % TC_STAGES, ph = (ks / TC_STAGES) & 1;
                mbar_wait(smem_u32(&full_bar[s]), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(sA + s * A_STAGE_BYTES), b0 = smem_u32(sB + s * B_STAGE_BYTES);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint64_t da = smem_desc(a0 + i * A_PLANE_BYTES, TC_BM * 16, 128);
#pragma unroll
                    for (int j0 = 0; j0 + i < 8; j0 += 4) {
                        const int planes = (8 - i - j0) < 4 ? (8 - i - j0) : 4;
                        const uint32_t n = (uint32_t)planes * TC_BN;
                        const uint64_t db = smem_desc(b0 + (uint32_t)j0 * TC_BN * 16, 8 * TC_BN * 16, 128);
                        tc_mma_i8(tmem_base + (uint32_t)(i + j0) * TC_BN, da, db, idesc0 | ((n >> 3) << 17), (ks > 0 || i > 0) ? 1u : 0u);
                    }
                }
                tc_commit_multicast(smem_u32(&empty_bar[s]), all);  // this CTA is done with stage s: tell all four producers
            }
            tc_commit(smem_u32(&acc_bar));
        }
    } else {
        mbar_wait(smem_u32(&acc_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t q = warp & 3;
        const uint32_t gm = mb * TC_BM + q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < TC_BN / 16; ++c) {
            u64 acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0;
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                uint32_t r[16];
                const uint32_t taddr = lane_addr + (uint32_t)(d * TC_BN + c * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += (u64)r[j] << (8 * d);
            }
            if (gm < a.M) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t gn = nb * TC_BN + c * 16 + j;
                    if (gn < a.N) {
                        const size_t o = (size_t)gm * a.N + gn;
                        u64 v = acc[j];
                        if (a.Z) v += a.Z[o];
                        if (a.accumulate) v += a.C[o];
                        a.C[o] = trunc_share(v, a.f, a.share);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();  // peers may still arrive on this CTA's empty barriers until all four are done
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}
#endif  // CGB_EXPERIMENTAL_TC_MC

}  // namespace

// One tensor-core launch over a K range that fits the 32-bit diagonal accumulators (n_pairs * K <= 4096).
static int tc_chunk(cgb_ctx* ctx, const u64* const A[2], const u64* const B[2], int n_pairs, const u64* Z, u64* C, uint32_t M, uint32_t K,
                    uint32_t lda, uint32_t N, int transA, int accumulate, int f, int share) {
    const uint32_t ks_pair = (K + TC_BK - 1) / TC_BK;
    const uint32_t n_ksteps = ks_pair * (uint32_t)n_pairs;
    CGB_REQUIRE(ctx, (uint64_t)n_ksteps * TC_BK <= 4096, "cgb_matmul_tc: K chunk too long for the 32-bit diagonal accumulators");
    const uint32_t Mpad = (M + TC_BM - 1) / TC_BM * TC_BM, Npad = (N + TC_BN - 1) / TC_BN * TC_BN;
    const size_t a_bytes = (size_t)(Mpad / TC_BM) * n_ksteps * A_STAGE_BYTES;
    const size_t b_bytes = (size_t)(Npad / TC_BN) * n_ksteps * B_STAGE_BYTES;
    // planes live in a grow-only buffer of the CONTEXT (ctx->scratch may hold V + F of the caller); an outgrown buffer is
    // retired until the context is destroyed, because a captured CUDA graph may still hold its address
    const size_t need = ((a_bytes + 255) & ~(size_t)255) + b_bytes;
    if (need > ctx->tc_planes_bytes) {
        if (ctx->tc_planes) ctx->retired.push_back(ctx->tc_planes);
        ctx->tc_planes = nullptr;
        ctx->tc_planes_bytes = 0;
        CGB_CHECK_CUDA(ctx, cudaMalloc(&ctx->tc_planes, need));
        ctx->tc_planes_bytes = need;
    }
    uint8_t* dA = (uint8_t*)ctx->tc_planes;
    uint8_t* dB = dA + ((a_bytes + 255) & ~(size_t)255);
    static const bool one_tile_per_cta = getenv("CGB_MATMUL_TC_V1") != nullptr;  // round-1 kernel, kept for A/B
    // B planes first (small), on the main stream
    for (int p = 0; p < n_pairs; ++p) {
        const uint64_t tb = (uint64_t)Npad * ks_pair * 2;
        limb_split_cols_kernel<<<(unsigned)((tb + 255) / 256), 256, 0, ctx->stream>>>(B[p], dB, K, N, Npad, p * ks_pair, n_ksteps, ks_pair);
        CGB_CHECK_LAUNCH(ctx, "limb_split_cols_kernel");
    }
    // Row groups (EXPERIMENT, off by default; CGB_MATMUL_TC_GROUPS=1): the limb split of A is HBM work, the tensor kernel is
    // tensor-pipe work, so the split of group g + 1 could run on an auxiliary stream beside the tensor kernel of group g.
    // Measured (profiles/r2k_matmul.json vs r2j_matmul.json): SLOWER -- 9.79 vs 8.28 ms at 2^20 x 512 x 512 -- the split's CTAs
    // take issue slots and L2 bandwidth from the persistent tensor CTAs and every group adds a kernel tail.  The split therefore
    // runs first, on the main stream (1.7 ms of the 8.3).
    const uint32_t n_mb_all = Mpad / TC_BM;
    uint32_t n_groups = 1;
    static const bool want_groups = getenv("CGB_MATMUL_TC_GROUPS") != nullptr;
    if (want_groups && !one_tile_per_cta && n_mb_all >= 8u * (uint32_t)ctx->num_sms)
        n_groups = std::min<uint32_t>(8, n_mb_all / (2u * (uint32_t)ctx->num_sms));
    if (n_groups > 1 && !ctx->tc_aux) {
        CGB_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->tc_aux, cudaStreamNonBlocking));
        for (int i = 0; i < 9; ++i) CGB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->tc_ev[i], cudaEventDisableTiming));
    }
    const uint32_t mb_per_group = (n_mb_all + n_groups - 1) / n_groups;
    auto split_rows = [&](cudaStream_t st, uint32_t mb0, uint32_t mb1) -> int {
        const uint32_t r0 = mb0 * TC_BM, rows = std::min(M, mb1 * TC_BM) - r0, rows_pad = (mb1 - mb0) * TC_BM;
        for (int p = 0; p < n_pairs; ++p) {
            const u64* src = transA ? A[p] + r0 : A[p] + (size_t)r0 * lda;
            const uint64_t ta = (uint64_t)rows_pad * ks_pair * 2;
            limb_split_rows_kernel<<<(unsigned)((ta + 255) / 256), 256, 0, st>>>(src, dA + (size_t)mb0 * n_ksteps * A_STAGE_BYTES, rows, K, lda,
                                                                                rows_pad, p * ks_pair, n_ksteps, ks_pair, transA);
            CGB_CHECK_LAUNCH(ctx, "limb_split_rows_kernel");
        }
        return CGB_OK;
    };
    if (n_groups > 1) {
        if (!ctx->tc_p_attr_set) {
            CGB_CHECK_CUDA(ctx, cudaFuncSetAttribute(matmul_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)TC_P_SMEM_BYTES));
            ctx->tc_p_attr_set = true;
        }
        // fork: the auxiliary stream starts after everything queued so far (the planes buffer may still be read by an earlier launch)
        CGB_CHECK_CUDA(ctx, cudaEventRecord(ctx->tc_ev[8], ctx->stream));
        CGB_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->tc_aux, ctx->tc_ev[8], 0));
        for (uint32_t g = 0; g < n_groups; ++g) {
            const uint32_t mb0 = g * mb_per_group, mb1 = std::min(n_mb_all, mb0 + mb_per_group);
            if (mb0 >= mb1) break;
            int rc = split_rows(ctx->tc_aux, mb0, mb1);
            if (rc) return rc;
            CGB_CHECK_CUDA(ctx, cudaEventRecord(ctx->tc_ev[g], ctx->tc_aux));
        }
        for (uint32_t g = 0; g < n_groups; ++g) {
            const uint32_t mb0 = g * mb_per_group, mb1 = std::min(n_mb_all, mb0 + mb_per_group);
            if (mb0 >= mb1) break;
            CGB_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->tc_ev[g], 0));  // join: group g's planes are written
            const uint32_t r0 = mb0 * TC_BM;
            TcPArgs pa;
            pa.t.A = dA + (size_t)mb0 * n_ksteps * A_STAGE_BYTES;
            pa.t.B = dB;
            pa.t.Z = Z ? Z + (size_t)r0 * N : nullptr;
            pa.t.C = C + (size_t)r0 * N;
            pa.t.M = std::min(M, mb1 * TC_BM) - r0;
            pa.t.N = N;
            pa.t.n_ksteps = n_ksteps;
            pa.t.f = f; pa.t.share = share; pa.t.accumulate = accumulate;
            pa.n_mb = mb1 - mb0;
            pa.n_nb = Npad / TC_BN;
            const unsigned ctas = (unsigned)std::min<uint32_t>(pa.n_mb * pa.n_nb, (uint32_t)ctx->num_sms);
            matmul_tc_persistent_kernel<<<ctas, 192, TC_P_SMEM_BYTES, ctx->stream>>>(pa);
            CGB_CHECK_LAUNCH(ctx, "matmul_tc_persistent_kernel");
        }
        ctx->last_kernel = "matmul_tc_persistent_kernel (tcgen05 kind::i8 limbs, TMEM diagonals, one CTA per SM; row groups, limb split of "
                           "the next group overlapped) + limb_split_{rows,cols}_kernel";
        return CGB_OK;
    }
    {
        int rc = split_rows(ctx->stream, 0, n_mb_all);
        if (rc) return rc;
    }
    if (!ctx->tc_attr_set) {  // a per-device attribute: set once per context, not once per process
        CGB_CHECK_CUDA(ctx, cudaFuncSetAttribute(matmul_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
        ctx->tc_attr_set = true;
    }
    TcArgs a;
    a.A = dA; a.B = dB; a.Z = Z; a.C = C; a.M = M; a.N = N; a.n_ksteps = n_ksteps;
    a.f = f; a.share = share; a.accumulate = accumulate;
    dim3 grid(Npad / TC_BN, Mpad / TC_BM);
#ifdef CGB_EXPERIMENTAL_TC_MC
    static const bool use_mc = getenv("CGB_MATMUL_IMPL") && std::string(getenv("CGB_MATMUL_IMPL")) == "tc_mc";
    if (use_mc && grid.x % TC_MC == 0) {  // whole clusters of four N-tiles only (N a multiple of 256)
        if (!ctx->tc_mc_attr_set) {
            CGB_CHECK_CUDA(ctx, cudaFuncSetAttribute(matmul_tc_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
            ctx->tc_mc_attr_set = true;
        }
        matmul_tc_mc_kernel<<<grid, 192, TC_SMEM_BYTES, ctx->stream>>>(a);
        CGB_CHECK_LAUNCH(ctx, "matmul_tc_mc_kernel");
        return CGB_OK;
    }
#endif
    if (!one_tile_per_cta) {
        if (!ctx->tc_p_attr_set) {
            CGB_CHECK_CUDA(ctx, cudaFuncSetAttribute(matmul_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)TC_P_SMEM_BYTES));
            ctx->tc_p_attr_set = true;
        }
        TcPArgs pa;
        pa.t = a;
        pa.n_mb = Mpad / TC_BM;
        pa.n_nb = Npad / TC_BN;
        const uint32_t n_tiles = pa.n_mb * pa.n_nb;
        const unsigned ctas = (unsigned)std::min<uint32_t>(n_tiles, (uint32_t)ctx->num_sms);
        matmul_tc_persistent_kernel<<<ctas, 192, TC_P_SMEM_BYTES, ctx->stream>>>(pa);
        CGB_CHECK_LAUNCH(ctx, "matmul_tc_persistent_kernel");
        ctx->last_kernel = "matmul_tc_persistent_kernel (tcgen05 kind::i8 limbs, TMEM diagonals, one CTA per SM) + limb_split_{rows,cols}_kernel";
        return CGB_OK;
    }
    matmul_tc_kernel<<<grid, 192, TC_SMEM_BYTES, ctx->stream>>>(a);
    CGB_CHECK_LAUNCH(ctx, "matmul_tc_kernel");
    ctx->last_kernel = "matmul_tc_kernel (tcgen05 kind::i8 limbs, TMEM diagonals) + limb_split_{rows,cols}_kernel";
    return CGB_OK;
}

// u8 x u8 limb MACs per second the tensor pipe sustains on the kernel's own MMA mix with free operands (synchronises)
extern "C" int cgb_probe_tensor_i8_peak(cgb_ctx* ctx, double* limb_mac_per_s) {
    CGB_REQUIRE(ctx, limb_mac_per_s, "cgb_probe_tensor_i8_peak: null argument");
    CGB_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t smem = A_STAGE_BYTES + B_STAGE_BYTES + 1024;
    CGB_CHECK_CUDA(ctx, cudaFuncSetAttribute(tc_pipe_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CGB_CHECK_CUDA(ctx, cudaEventCreate(&e0));
    CGB_CHECK_CUDA(ctx, cudaEventCreate(&e1));
    const uint32_t n_ksteps = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CGB_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        tc_pipe_probe_kernel<<<ctx->num_sms, 64, smem, ctx->stream>>>(n_ksteps);
        CGB_CHECK_LAUNCH(ctx, "tc_pipe_probe_kernel");
        CGB_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        CGB_CHECK_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        CGB_CHECK_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double macs = (double)ctx->num_sms * n_ksteps * 36.0 * TC_BM * TC_BN * TC_BK;
        if (rep > 0 && ms > 0.f) best = std::max(best, macs / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *limb_mac_per_s = best;
    return CGB_OK;
}

// Tensor-core product of up to two (A, B) pairs accumulated into one result (the Beaver finish E*(V[+F]) + U*F), with the
// same epilogue options as the integer-pipe kernel.  K is processed in chunks that keep the diagonal accumulators exact;
// the partial results are carried in C (u64), Z is added with the first chunk and the truncation applied with the last.
int cgb_matmul_tc_run(cgb_ctx* ctx, const u64* const A[2], const u64* const B[2], int n_pairs, const u64* Z, u64* C, uint32_t M,
                      uint32_t K, uint32_t N, int transA, int accumulate, int f, int share) {
    const uint32_t kmax = (4096u / (uint32_t)n_pairs) / TC_BK * TC_BK;
    const uint32_t lda = transA ? M : K;
    for (uint32_t k0 = 0; k0 < K || k0 == 0; k0 += kmax) {
        const uint32_t kc = std::min(kmax, K - k0);
        const bool first = k0 == 0, last = k0 + kc >= K;
        const u64* Ac[2] = {nullptr, nullptr};
        const u64* Bc[2] = {nullptr, nullptr};
        for (int p = 0; p < n_pairs; ++p) {
            Ac[p] = transA ? A[p] + (size_t)k0 * M : A[p] + k0;
            Bc[p] = B[p] + (size_t)k0 * N;
        }
        int rc = tc_chunk(ctx, Ac, Bc, n_pairs, first ? Z : nullptr, C, M, kc, lda, N, transA, first ? accumulate : 1, last ? f : 0, share);
        if (rc) return rc;
        if (last) break;
    }
    return CGB_OK;
}
