// common.cuh -- shared definitions for the sm_100a kernels behind include/cognn_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/cognn_b200.h"

struct cgb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    std::string err;
    uint64_t launches = 0;
    int matmul_impl = -1;  // -1: CGB_MATMUL_IMPL / auto; 0 auto, 1 integer pipe, 2 tensor pipe (cgb_ctx_set_matmul_impl)
    const char* last_kernel = "";  // name of the gather / matmul kernel the last dispatch chose (bench.py reports it)
    // device word added to the stream id of every PRG launch (cgb_ctx_set_prg_stream_bias): lets a captured CUDA graph
    // draw fresh randomness at every replay
    const uint64_t* prg_bias = nullptr;
    // grow-only device scratch (split-K accumulators, Beaver temporaries)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // limb planes of the tensor-core matmul (grow-only, per context = per device and stream) and its per-device attribute flags
    void* tc_planes = nullptr;
    size_t tc_planes_bytes = 0;
    bool tc_attr_set = false, tc_mc_attr_set = false, tc_p_attr_set = false;
    cudaStream_t tc_aux = nullptr;  // limb split of the next row group runs here, beside the tensor kernel of the current one
    cudaEvent_t tc_ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::vector<void*> retired;  // outgrown buffers, freed with the context (captured graphs may still reference them)
    // double-buffered staging for the pipelined host entry point (cgb_host_gather_sum_async)
    struct HostPipe {
        cudaStream_t h2d = nullptr, d2h = nullptr;
        cudaEvent_t ev_x[2] = {nullptr, nullptr}, ev_y[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
        void* buf[2] = {nullptr, nullptr};
        size_t bytes[2] = {0, 0};
        uint64_t steps = 0;
        bool ready = false;
    } pipe;
};

struct cgb_csr {
    uint32_t n_rows = 0;
    uint32_t n_src_rows = 0;
    uint64_t n_edges = 0;
    uint32_t* d_rowptr = nullptr;  // n_rows + 1
    uint32_t* d_col = nullptr;     // n_edges
    // rows with more than CGB_LONG_ROW edges are cut into slices of CGB_SLICE_EDGES edges
    uint32_t n_long_rows = 0;
    uint32_t n_slices = 0;
    uint32_t* d_slice_row = nullptr;    // n_slices: destination row
    uint32_t* d_slice_begin = nullptr;  // n_slices: first edge
    uint32_t* d_slice_first = nullptr;  // n_slices: index of the row's first slice
    uint32_t* d_slice_count = nullptr;  // n_slices: number of slices of the row
    uint32_t* d_long_id = nullptr;      // n_slices: dense id of the long row (counter index)
    uint32_t* d_counters = nullptr;     // n_long_rows * max_col_tiles arrival counters (self-resetting)
    uint32_t counters_len = 0;
    uint64_t* d_partial = nullptr;      // n_slices x D partial sums (grow-only)
    size_t partial_words = 0;
    // ---- edge-balanced schedule (gather_chunk_kernel): fixed chunks of 1 << chunk_shift edges ----
    uint32_t* d_colf = nullptr;         // col | CGB_END_FLAG on the last edge of every row
    uint32_t* d_nz_row = nullptr;       // ids of the non-empty rows, ascending
    uint32_t* d_empty_row = nullptr;    // ids of the empty rows
    uint32_t* d_chunk_nz = nullptr;     // per chunk: index into nz_row of the row holding its first edge;
                                        // bit 31: that row started in an earlier chunk
    uint32_t n_nz = 0, n_empty = 0, n_chunks = 0;
    uint32_t chunk_shift = 6;           // log2(edges per chunk): 64 edges, 128 from CGB_BIG_GRAPH_EDGES edges on
    uint32_t* d_chunk_ctr = nullptr;    // n_chunks * col tiles arrival counters (self-resetting)
    uint32_t chunk_ctr_len = 0;
    uint64_t* d_piece = nullptr;        // 2 * n_chunks x D: [head pieces | tail pieces] (grow-only)
    size_t piece_words = 0;
    std::vector<void*> retired;         // outgrown scratch buffers, freed with the handle (graphs may still reference them)
    std::vector<uint32_t> h_nz_row;     // host copy of d_nz_row (block bookkeeping of cgb_gather_sum_signal)
    uint32_t* d_sig_done = nullptr;     // per-block completion counters of cgb_gather_sum_signal (self-resetting)
};
// edges per chunk of the edge-balanced schedule = 1 << chunk_shift, chosen per CSR when it is built: 64 keeps small graphs
// (one wave of groups or less) short, 128 halves the per-chunk prologues and boundary pieces of big ones (100M edges,
// D = 16: 1.77 ms against 1.88 ms); CGB_CHUNK_SHIFT=6|7|8 overrides
#define CGB_CHUNK_SHIFT_SMALL 6u
#define CGB_CHUNK_SHIFT_BIG 7u
#define CGB_BIG_GRAPH_EDGES (1ull << 24)
#define CGB_MAX_BLOCKS 16
#define CGB_END_FLAG 0x80000000u

#define CGB_LONG_ROW 256u
#define CGB_SLICE_EDGES 256u

static inline int cgb_fail(cgb_ctx* ctx, int code, const char* what, const char* detail) {
    if (ctx) {
        ctx->err = std::string(what) + ": " + (detail ? detail : "");
    }
    return code;
}

#define CGB_CHECK_CUDA(ctx, call)                                                   \
    do {                                                                            \
        cudaError_t _e = (call);                                                    \
        if (_e != cudaSuccess) return cgb_fail((ctx), CGB_ERR_CUDA, #call, cudaGetErrorString(_e)); \
    } while (0)

#define CGB_CHECK_LAUNCH(ctx, name)                                                 \
    do {                                                                            \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) return cgb_fail((ctx), CGB_ERR_CUDA, name, cudaGetErrorString(_e)); \
        (ctx)->launches++;                                                          \
    } while (0)

#define CGB_REQUIRE(ctx, cond, msg)                                                 \
    do {                                                                            \
        if (!(cond)) return cgb_fail((ctx), CGB_ERR_INVALID, msg, #cond);           \
    } while (0)

int cgb_scratch_reserve(cgb_ctx* ctx, size_t bytes);

// ---- device helpers ----------------------------------------------------------------------------------------
typedef unsigned long long u64;

__device__ __forceinline__ u64 trunc_share(u64 z, int f, int share) {
    // SecureML local truncation (DESIGN.md "Frozen semantics"): share 0 logical shift, share 1 negate-shift-negate
    if (f <= 0) return z;
    return share == 0 ? (z >> f) : (0ull - ((0ull - z) >> f));
}

// 128-bit read-only gather load, no L1 allocation (rows are re-used only through L2)
__device__ __forceinline__ ulonglong2 ld_nc_v2(const u64* p) {
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ u64 ld_nc_u64(const u64* p) {
    u64 r;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
// L2-coherent load (partials written by other CTAs in the same launch)
__device__ __forceinline__ u64 ld_cg_u64(const u64* p) {
    u64 r;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");  // ordered after the __threadfence before it
    return r;
}
__device__ __forceinline__ void st_cs_v2(u64* p, u64 a, u64 b) {
    asm volatile("st.global.cs.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_cs_u64(u64* p, u64 a) {
    asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
}
