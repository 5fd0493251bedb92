// gather.cu -- op (1): edge-wise scatter / gather-sum of Z_2^64 share rows (CSR-by-destination SpMM).
//
// Serves: client_oblivious_mapper_online (ss_vertex_centric_algo_kernel.h:752,760,
// 818,848), ScatterComp copy (optimize-gcn/gcn.h:300) and prefix_network_aggregate ADD_AGG (gcn.h:328-335).
//
// Layout: share rows are dense row-major u64, one row = D columns.  A "group" of LANES consecutive lanes owns one
// destination row (x one column tile of VEC*LANES columns); each lane keeps VEC accumulators and issues U
// independent 8*VEC-byte gather loads per step, so a warp has 32*U loads in flight.  Rows longer than
// CGB_LONG_ROW edges are cut (once, at cgb_csr_create) into slices of CGB_SLICE_EDGES edges that are summed by
// separate groups; the last slice to arrive (self-resetting arrival counter) folds the partial sums, so a
// power-law tail cannot serialise on one group.  Integer addition mod 2^64 is associative and commutative, so
// every schedule yields the same bits.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

// 256-bit row accesses by default (round 1b: +22 % at D = 16, +34 % at D = 64, +42 % at D = 128 over the 128-bit form,
// profiles/r1b_sweep_gather_vec*.jsonl); CGB_GATHER_VEC=2 selects the 128-bit kernels for A/B runs
#ifndef CGB_GATHER_DEFAULT_VEC
#define CGB_GATHER_DEFAULT_VEC 4
#endif

namespace {

struct GatherArgs {
    const uint32_t* rowptr;
    const uint32_t* col;
    const u64* x;
    const u64* delta;
    u64* y;
    uint32_t n_rows;
    uint32_t D;
    uint32_t n_ct;  // column tiles per row
    // long-row slices
    uint32_t n_slices;
    const uint32_t* slice_row;
    const uint32_t* slice_begin;
    const uint32_t* slice_first;
    const uint32_t* slice_count;
    const uint32_t* long_id;
    uint32_t* counters;
    u64* partial;
};

template <int VEC>
struct Acc;
template <>
struct Acc<1> {
    u64 a;
    __device__ __forceinline__ void zero() { a = 0; }
    __device__ __forceinline__ void load_nc(const u64* p) { a = ld_nc_u64(p); }
    __device__ __forceinline__ void load_cg(const u64* p) { a = ld_cg_u64(p); }
    __device__ __forceinline__ void add(const Acc& o) { a += o.a; }
    __device__ __forceinline__ void store(u64* p) const { *p = a; }
    __device__ __forceinline__ void store_cs(u64* p) const { st_cs_u64(p, a); }
};
template <>
struct Acc<2> {
    u64 a, b;
    __device__ __forceinline__ void zero() { a = 0; b = 0; }
    __device__ __forceinline__ void load_nc(const u64* p) {
        ulonglong2 v = ld_nc_v2(p);
        a = v.x; b = v.y;
    }
    __device__ __forceinline__ void load_cg(const u64* p) {
        a = ld_cg_u64(p); b = ld_cg_u64(p + 1);
    }
    __device__ __forceinline__ void add(const Acc& o) { a += o.a; b += o.b; }
    __device__ __forceinline__ void store(u64* p) const { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b); }
    __device__ __forceinline__ void store_cs(u64* p) const { st_cs_v2(p, a, b); }
};

// 256-bit accesses (LDG.E.256 / STG.E.256, new on sm_100): four u64 columns per lane, so a D = 16 row is one load
// instruction of four lanes -- half the address arithmetic, shuffles and flag tests per byte of the 128-bit form
template <>
struct Acc<4> {
    u64 a, b, c, d;
    __device__ __forceinline__ void zero() { a = 0; b = 0; c = 0; d = 0; }
    __device__ __forceinline__ void load_nc(const u64* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    }
    // the piece paths (rows cut by a chunk boundary) are rare: plain 64-/128-bit accesses there
    __device__ __forceinline__ void load_cg(const u64* p) {
        a = ld_cg_u64(p); b = ld_cg_u64(p + 1); c = ld_cg_u64(p + 2); d = ld_cg_u64(p + 3);
    }
    __device__ __forceinline__ void add(const Acc& o) { a += o.a; b += o.b; c += o.c; d += o.d; }
    __device__ __forceinline__ void store(u64* p) const {
        reinterpret_cast<ulonglong2*>(p)[0] = make_ulonglong2(a, b);
        reinterpret_cast<ulonglong2*>(p)[1] = make_ulonglong2(c, d);
    }
    __device__ __forceinline__ void store_cs(u64* p) const {
        asm volatile("st.global.cs.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
    }
};

template <int LANES>
__device__ __forceinline__ unsigned group_mask(int lane_in_warp) {
    if constexpr (LANES == 32) return 0xffffffffu;
    else return ((1u << LANES) - 1u) << (lane_in_warp & ~(LANES - 1));
}

// sum of x[col[e], col0 .. col0+VEC) for e in [e0, e1)
template <int VEC, int LANES, int U, bool IDENT>
__device__ __forceinline__ Acc<VEC> accumulate_edges(const uint32_t* __restrict__ col, const u64* __restrict__ x,
                                                     uint32_t e0, uint32_t e1, uint32_t D, uint32_t col0, bool active,
                                                     int lane, unsigned mask) {
    Acc<VEC> acc;
    acc.zero();
    for (uint32_t e = e0; e < e1; e += LANES) {
        const uint32_t n = min((uint32_t)LANES, e1 - e);
        uint32_t my = 0;
        if (IDENT) my = e + lane;
        else if ((uint32_t)lane < n) my = __ldg(col + e + lane);
        for (uint32_t k0 = 0; k0 < n; k0 += U) {
            Acc<VEC> v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t idx = __shfl_sync(mask, my, k0 + u, LANES);
                if (active && (k0 + u < n)) v[u].load_nc(x + (size_t)idx * D + col0);
                else v[u].zero();
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc.add(v[u]);
        }
    }
    return acc;
}

template <int VEC, int LANES, int U>
__global__ void __launch_bounds__(256) gather_sum_kernel(const GatherArgs a) {
    constexpr int GROUPS = 256 / LANES;
    const int lane = threadIdx.x & (LANES - 1);
    const unsigned mask = group_mask<LANES>(threadIdx.x & 31);
    const uint64_t gid = (uint64_t)blockIdx.x * GROUPS + threadIdx.x / LANES;
    const uint64_t total = ((uint64_t)a.n_slices + a.n_rows) * a.n_ct;
    if (gid >= total) return;
    const uint32_t item = (uint32_t)(gid / a.n_ct);
    const uint32_t ct = (uint32_t)(gid - (uint64_t)item * a.n_ct);
    const uint32_t col0 = ct * (VEC * LANES) + lane * VEC;
    const bool active = col0 < a.D;

    if (item >= a.n_slices) {
        // ---- one whole destination row ----
        const uint32_t row = item - a.n_slices;
        const uint32_t b = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
        if (e - b > CGB_LONG_ROW) return;  // summed by its slices
        Acc<VEC> acc = accumulate_edges<VEC, LANES, U, false>(a.col, a.x, b, e, a.D, col0, active, lane, mask);
        if (active) {
            const size_t o = (size_t)row * a.D + col0;
            if (a.delta) {
                Acc<VEC> d;
                d.load_nc(a.delta + o);
                acc.add(d);
            }
            acc.store_cs(a.y + o);
        }
    } else {
        // ---- one slice of a long row ----
        const uint32_t s = item;
        const uint32_t row = __ldg(a.slice_row + s);
        const uint32_t b = __ldg(a.slice_begin + s);
        const uint32_t e = min(b + CGB_SLICE_EDGES, __ldg(a.rowptr + row + 1));
        Acc<VEC> acc = accumulate_edges<VEC, LANES, U, false>(a.col, a.x, b, e, a.D, col0, active, lane, mask);
        if (active) acc.store(a.partial + (size_t)s * a.D + col0);
        __threadfence();
        __syncwarp(mask);
        const uint32_t cnt = __ldg(a.slice_count + s);
        uint32_t* ctr = a.counters + (size_t)__ldg(a.long_id + s) * a.n_ct + ct;
        uint32_t prev = 0;
        if (lane == 0) prev = atomicAdd(ctr, 1u);
        prev = __shfl_sync(mask, prev, 0, LANES);
        if (prev == cnt - 1) {
            __threadfence();
            if (active) {
                const uint32_t first = __ldg(a.slice_first + s);
                const size_t o = (size_t)row * a.D + col0;
                Acc<VEC> sum;
                sum.zero();
                if (a.delta) sum.load_nc(a.delta + o);
                for (uint32_t k = 0; k < cnt; ++k) {
                    Acc<VEC> p;
                    p.load_cg(a.partial + (size_t)(first + k) * a.D + col0);
                    sum.add(p);
                }
                sum.store_cs(a.y + o);
            }
            if (lane == 0) *ctr = 0;  // ready for the next launch
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// Edge-balanced schedule.  The row-per-group kernel above leaves a power-law graph latency bound: a block lives
// as long as its longest row, and every row pays the rowptr -> col -> x dependent-load chain (ncu, round 1:
// 25% DRAM utilisation, 34% of peak warps active).  Here every group owns exactly 1 << chunk_shift (64 or 128) consecutive
// edges, whatever rows they belong to.  Row boundaries travel in bit 31 of the column index (set on the last
// edge of each row), the next batch of indices is prefetched while the current batch of row loads is in flight,
// rows that lie inside one chunk are stored directly, and a row cut by a chunk boundary leaves one "piece" per
// chunk that the last arriving chunk folds (arrival counter indexed by the row's first chunk, self-resetting).
// ------------------------------------------------------------------------------------------------------------
struct ChunkArgs {
    const uint32_t* colf;
    const uint32_t* chunk_nz;
    const uint32_t* nz_row;
    const uint32_t* rowptr;
    const uint32_t* empty_row;
    const u64* x;
    const u64* delta;
    u64* y;
    uint32_t n_chunks, n_empty, n_edges;
    uint32_t chunk_shift;  // log2(edges per chunk)
    uint32_t D, n_ct;
    uint32_t* counters;
    u64* piece_head;  // n_chunks x D: sum of the chunk's leading edges when they continue a row from an earlier chunk
    u64* piece_tail;  // n_chunks x D: sum of the chunk's trailing edges when their row continues in a later chunk
    // optional: the output rows are split into n_blk contiguous blocks, block t stored at blk_base[t] -- the base may
    // be PEER memory (another GPU's receive buffer mapped over NVLink), which fuses the mirror-update exchange of
    // ssk.h:835 -> 1067/1090 into the gather: every 8*D-byte row is written exactly once, straight to its consumer
    int n_blk;
    u64* blk_base[CGB_MAX_BLOCKS];
    uint32_t blk_off[CGB_MAX_BLOCKS + 1];
    // compact output: row k of y is the k-th NON-EMPTY destination row (nz_row[k]); rows without edges are not written at
    // all.  This is the mirror-update block a party sends to another party: which destinations of the receiver have an edge
    // from the sender is public to both (sendPosVec / recvPosVec, ssk.h:507-516), so empty rows need not cross NVLink
    uint32_t compact;
};

__device__ __forceinline__ u64* out_row(const ChunkArgs& a, uint32_t row) {
    if (a.n_blk == 0) return a.y + (size_t)row * a.D;
    int t = 0;
#pragma unroll 1
    while (t + 1 < a.n_blk && row >= a.blk_off[t + 1]) ++t;
    return a.blk_base[t] + (size_t)(row - a.blk_off[t]) * a.D;
}

// ---- signalling mode (cgb_gather_sum_signal) ---------------------------------------------------------------------------
// The output rows are cut into up to CGB_MAX_SIG contiguous blocks (mirror-update blocks for the other parties, or pieces of
// them, and the party's own block).  Each block has its own destination buffer, dense (row - first row of the block) or compact
// (position among the block's non-empty rows), and a FLAG: every CTA counts the rows it completes per block, and the CTA
// that brings a block's count to its total raises the flag (st.release.sys, usually into the CONSUMER's memory over NVLink)
// while the rest of the grid is still gathering the later blocks.  The transfer of block t (the consumer pulls it as soon
// as it sees the flag) therefore overlaps the gather of blocks t+1.. inside ONE launch.
#define CGB_MAX_SIG 32
struct SigArgs {
    uint32_t n_sig;
    uint32_t compact_mask;            // bit b: block b is stored in compact form
    uint32_t off[CGB_MAX_SIG + 1];    // first row of each block (ascending), off[n_sig] = n_rows
    uint32_t k0[CGB_MAX_SIG];         // number of non-empty rows before the block
    uint32_t total[CGB_MAX_SIG];      // (row, column tile) stores that complete the block
    u64* base[CGB_MAX_SIG];
    uint32_t* flag[CGB_MAX_SIG];      // may be null (no signal), may be peer memory
    uint32_t* done;                   // CGB_MAX_SIG device counters, self-resetting
    uint32_t value;
    // rows without edges exist in the output of DENSE blocks only: e_cum[b] = number of such rows in the dense blocks before b
    // (e_cum[n_sig] = all of them); the zero rows of block b are empty_row[off[b] - k0[b] + i], i < e_cum[b + 1] - e_cum[b]
    uint32_t e_cum[CGB_MAX_SIG + 1];
};

__device__ __forceinline__ uint32_t sig_block_of(const SigArgs& g, uint32_t row, uint32_t b) {
#pragma unroll 1
    while (b + 1 < g.n_sig && row >= g.off[b + 1]) ++b;
    return b;
}
__device__ __forceinline__ u64* sig_row_ptr(const SigArgs& g, uint32_t b, uint32_t row, uint32_t k, uint32_t D) {
    const uint32_t idx = ((g.compact_mask >> b) & 1u) ? (k - g.k0[b]) : (row - g.off[b]);
    return g.base[b] + (size_t)idx * D;
}

// A row segment cut by a chunk boundary has been stored as a "piece"; count the arrival and, if this was the last
// piece of the row, fold them.  `orow` is the row's output index (the row itself, or its position among the non-empty rows in
// compact mode).
template <int VEC, int LANES, int MODE = 0>
__device__ __forceinline__ void piece_arrive_body(const ChunkArgs& a, uint32_t row, uint32_t orow, uint32_t ct, uint32_t col0,
                                                  bool active, int lane, unsigned mask, const SigArgs* g = nullptr,
                                                  uint32_t* s_cnt = nullptr) {
    __threadfence();
    __syncwarp(mask);
    const uint32_t rb = __ldg(a.rowptr + row), re = __ldg(a.rowptr + row + 1);
    const uint32_t c1 = rb >> a.chunk_shift, c2 = (re - 1) >> a.chunk_shift;
    uint32_t* ctr = a.counters + (size_t)c1 * a.n_ct + ct;
    uint32_t prev = 0;
    if (lane == 0) prev = atomicAdd(ctr, 1u);
    prev = __shfl_sync(mask, prev, 0, LANES);
    if (prev == c2 - c1) {  // last of the c2 - c1 + 1 pieces
        __threadfence();
        if (active) {
            const size_t o = (size_t)orow * a.D + col0;
            Acc<VEC> sum;
            sum.zero();
            if (a.delta) sum.load_nc(a.delta + o);
            Acc<VEC> p;
            p.load_cg(a.piece_tail + (size_t)c1 * a.D + col0);
            sum.add(p);
            for (uint32_t cc = c1 + 1; cc <= c2; ++cc) {
                p.load_cg(a.piece_head + (size_t)cc * a.D + col0);
                sum.add(p);
            }
            if constexpr (MODE == 2) {  // orow carries the row's index among the non-empty rows
                const uint32_t b = sig_block_of(*g, row, 0);
                sum.store_cs(sig_row_ptr(*g, b, row, orow, a.D) + col0);
            } else {
                sum.store_cs(out_row(a, orow) + col0);
            }
        }
        if constexpr (MODE == 2) {
            if (lane == 0) atomicAdd(&s_cnt[sig_block_of(*g, row, 0)], 1u);
        }
        if (lane == 0) *ctr = 0;  // ready for the next launch
    }
}

// out-of-line form for the cp.async variant (kept for A/B runs)
template <int VEC, int LANES>
__device__ __noinline__ void piece_arrive(const ChunkArgs& a, uint32_t row, uint32_t ct, uint32_t col0, bool active, int lane,
                                          unsigned mask) {
    piece_arrive_body<VEC, LANES>(a, row, row, ct, col0, active, lane, mask);
}

// IPL: column indices each lane holds per batch (a batch is LANES * IPL edges; narrow 256-bit groups of 2 or 4 lanes keep
// 8 edges per batch this way, so U can stay above the lane count)
//
// MODE 0: rows stored at their index (or into the blk_* windows); 1: compact (COMPACT); 2: signalling blocks (SigArgs)
template <int VEC, int LANES, int U, int BLOCK, int IPL, int MODE>
__device__ __forceinline__ void gather_chunk_body(const ChunkArgs& a, const SigArgs* g, uint32_t* s_cnt) {
    constexpr bool COMPACT = MODE == 1;
    constexpr int BATCH = LANES * IPL;
    constexpr int GROUPS = BLOCK / LANES;
    const int lane = threadIdx.x & (LANES - 1);
    const unsigned mask = group_mask<LANES>(threadIdx.x & 31);
    const uint64_t gid = (uint64_t)blockIdx.x * GROUPS + threadIdx.x / LANES;
    uint32_t n_empty_items = COMPACT ? 0u : a.n_empty;
    if constexpr (MODE == 2) n_empty_items = g->e_cum[g->n_sig];  // only the dense blocks have rows without edges to write
    const uint64_t total = ((uint64_t)a.n_chunks + n_empty_items) * a.n_ct;
    if (gid >= total) return;
    uint32_t item = (uint32_t)(gid / a.n_ct);
    const uint32_t ct = (uint32_t)(gid - (uint64_t)item * a.n_ct);
    const uint32_t col0 = ct * (VEC * LANES) + lane * VEC;
    const bool active = col0 < a.D;
    if constexpr (MODE == 2) {
        // the zero rows come FIRST in this mode (the first CTAs of the grid): a dense block is complete -- and its flag raised --
        // when its last chunk is, not when the tail of the grid gets to its zero rows
        if (item < n_empty_items) {
            uint32_t b = 0;
#pragma unroll 1
            while (item >= g->e_cum[b + 1]) ++b;
            const uint32_t row = __ldg(a.empty_row + (g->off[b] - g->k0[b]) + (item - g->e_cum[b]));
            if (active) {
                Acc<VEC> v;
                v.zero();
                v.store_cs(sig_row_ptr(*g, b, row, 0, a.D) + col0);
            }
            if (lane == 0) atomicAdd(&s_cnt[b], 1u);
            return;
        }
        item -= n_empty_items;
    }

    if (item >= a.n_chunks) {  // a row without edges: y = delta (or 0)
        const uint32_t row = __ldg(a.empty_row + (item - a.n_chunks));
        if (active) {
            const size_t o = (size_t)row * a.D + col0;
            Acc<VEC> v;
            v.zero();
            if (a.delta) v.load_nc(a.delta + o);
            v.store_cs(out_row(a, row) + col0);
        }
        return;
    }
    const uint32_t c = item;
    uint32_t e = c << a.chunk_shift;
    const uint32_t end = min(a.n_edges, e + (1u << a.chunk_shift));
    const uint32_t cn = __ldg(a.chunk_nz + c);
    uint32_t k = cn & ~CGB_END_FLAG;
    bool head_open = (cn & CGB_END_FLAG) != 0;
    bool open = false, have_head = false;
    uint32_t head_row = 0, head_k = 0;
    uint32_t sb = 0;  // MODE 2: block of the last row stored (rows ascend inside a chunk)
    Acc<VEC> acc;
    acc.zero();
    uint32_t my[IPL], nxt[IPL];
#pragma unroll
    for (int i = 0; i < IPL; ++i) my[i] = (e + i * LANES + lane < end) ? __ldg(a.colf + e + i * LANES + lane) : 0u;
    while (e < end) {
        const uint32_t n = min((uint32_t)BATCH, end - e);
#pragma unroll
        for (int i = 0; i < IPL; ++i)  // prefetch the next batch of indices
            nxt[i] = (e + BATCH + i * LANES + lane < end) ? __ldg(a.colf + e + BATCH + i * LANES + lane) : 0u;
        auto sub_batch = [&](const uint32_t k0) {
            Acc<VEC> v[U];
            uint32_t id[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                id[u] = __shfl_sync(mask, my[(k0 + u) / LANES], (k0 + u) % LANES, LANES);
                if (active && (k0 + u < n)) v[u].load_nc(a.x + (size_t)(id[u] & ~CGB_END_FLAG) * a.D + col0);
                else v[u].zero();
            }
            uint32_t any = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) any |= id[u];
            if (!(any & CGB_END_FLAG)) {  // no row ends inside this batch
#pragma unroll
                for (int u = 0; u < U; ++u) acc.add(v[u]);
                open = true;
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (k0 + u < n) {
                        acc.add(v[u]);
                        open = true;
                        if (id[u] & CGB_END_FLAG) {
                            const uint32_t row = __ldg(a.nz_row + k);
                            if (!head_open) {  // the row lies inside this chunk: store it
                                if constexpr (MODE == 2) {
                                    sb = sig_block_of(*g, row, sb);
                                    if (active) acc.store_cs(sig_row_ptr(*g, sb, row, k, a.D) + col0);
                                    if (lane == 0) atomicAdd(&s_cnt[sb], 1u);
                                } else if (active) {
                                    const uint32_t orow = COMPACT ? k : row;
                                    const size_t o = (size_t)orow * a.D + col0;
                                    if (a.delta) {
                                        Acc<VEC> d;
                                        d.load_nc(a.delta + o);
                                        acc.add(d);
                                    }
                                    acc.store_cs(out_row(a, orow) + col0);
                                }
                            } else {  // end of a row that began in an earlier chunk: leave a head piece (arrival counted below)
                                if (active) acc.store(a.piece_head + (size_t)c * a.D + col0);
                                head_row = row;
                                if constexpr (MODE != 0) head_k = k;
                                have_head = true;
                            }
                            acc.zero();
                            head_open = false;
                            open = false;
                            ++k;
                        }
                    }
                }
            }
        };
        if constexpr (IPL == 1) {  // the register-lean form the 128-bit kernels were tuned with (no unrolling over k0)
            for (uint32_t k0 = 0; k0 < n; k0 += U) sub_batch(k0);
        } else {  // unrolled, so that my[(k0 + u) / LANES] is a register, not a local-memory array
#pragma unroll
            for (uint32_t k0 = 0; k0 < (uint32_t)BATCH; k0 += U) {
                if (k0 >= n) break;
                sub_batch(k0);
            }
        }
        e += BATCH;
#pragma unroll
        for (int i = 0; i < IPL; ++i) my[i] = nxt[i];
    }
    // Piece arrivals after the edge loop, always INLINED (two copies): round 1 found that with the fold out of line (a call)
    // the IPL = 2 instantiation returned wrong columns 1..3 for rows cut by a chunk boundary -- the accumulator is live across
    // the first call.  No instantiation of this kernel calls out of line any more.
    if (have_head) piece_arrive_body<VEC, LANES, MODE>(a, head_row, MODE != 0 ? head_k : head_row, ct, col0, active, lane, mask, g, s_cnt);
    if (open) {  // the chunk ends inside a row
        if (active) acc.store((head_open ? a.piece_head : a.piece_tail) + (size_t)c * a.D + col0);
        const uint32_t row = __ldg(a.nz_row + k);
        piece_arrive_body<VEC, LANES, MODE>(a, row, MODE != 0 ? k : row, ct, col0, active, lane, mask, g, s_cnt);
    }
}

template <int VEC, int LANES, int U, int BLOCK, int OCC_THREADS, int IPL = 1, bool COMPACT = false>
__global__ void __launch_bounds__(BLOCK, OCC_THREADS / BLOCK)
gather_chunk_kernel(const __grid_constant__ ChunkArgs a) {
    gather_chunk_body<VEC, LANES, U, BLOCK, IPL, COMPACT ? 1 : 0>(a, nullptr, nullptr);
}

// The fused compute + signalling form: the same edge loop, then every CTA adds its per-block row counts to the global counters
// and the one that completes a block raises that block's flag.
template <int VEC, int LANES, int U, int BLOCK, int OCC_THREADS, int IPL = 1>
__global__ void __launch_bounds__(BLOCK, OCC_THREADS / BLOCK)
gather_chunk_signal_kernel(const __grid_constant__ ChunkArgs a, const __grid_constant__ SigArgs g) {
    __shared__ uint32_t s_cnt[CGB_MAX_SIG];
    if (threadIdx.x < CGB_MAX_SIG) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    gather_chunk_body<VEC, LANES, U, BLOCK, IPL, 2>(a, &g, s_cnt);
    __threadfence();  // this thread's rows are visible device-wide before its CTA reports them
    __syncthreads();
    if (threadIdx.x < g.n_sig) {
        const uint32_t b = threadIdx.x, cnt = s_cnt[b];
        if (cnt) {
            __threadfence();
            const uint32_t prev = atomicAdd(g.done + b, cnt);
            if (prev + cnt == g.total[b]) {  // this CTA completed block b
                atomicExch(g.done + b, 0u);  // ready for the next launch
                if (g.flag[b]) {
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(g.flag[b]), "r"(g.value) : "memory");
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// Shared-memory staged variant of the edge-balanced schedule.  ncu on gather_chunk_kernel (round 1): 42% DRAM
// utilisation, long-scoreboard stalls dominate, and occupancy is capped by the registers that hold loads in
// flight.  Here the row loads are cp.async (LDGSTS) copies into a double-buffered shared-memory ring -- every
// thread copies and later reads only its own 16-byte column slot, so no barrier is needed -- which moves the
// bytes in flight from the register file (about 80 KB/SM) to shared memory (about 220 KB/SM).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t smem_addr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int VEC>
__device__ __forceinline__ Acc<VEC> lds_acc(uint32_t smem_addr);
template <>
__device__ __forceinline__ Acc<2> lds_acc<2>(uint32_t smem_addr) {
    Acc<2> r;
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(r.a), "=l"(r.b) : "r"(smem_addr));
    return r;
}
template <>
__device__ __forceinline__ Acc<1> lds_acc<1>(uint32_t smem_addr) {
    Acc<1> r;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(r.a) : "r"(smem_addr));
    return r;
}

template <int VEC, int LANES, int U, int BLOCK, int NBUF>
__global__ void __launch_bounds__(BLOCK) gather_chunk_async_kernel(const __grid_constant__ ChunkArgs a) {
    constexpr int GROUPS = BLOCK / LANES;
    constexpr uint32_t SLOT = VEC * 8;  // bytes per thread per staged edge
    extern __shared__ __align__(16) unsigned char stage_raw[];
    const int lane = threadIdx.x & (LANES - 1);
    const unsigned mask = group_mask<LANES>(threadIdx.x & 31);
    const uint64_t gid = (uint64_t)blockIdx.x * GROUPS + threadIdx.x / LANES;
    const uint64_t total = ((uint64_t)a.n_chunks + a.n_empty) * a.n_ct;
    if (gid >= total) return;
    const uint32_t item = (uint32_t)(gid / a.n_ct);
    const uint32_t ct = (uint32_t)(gid - (uint64_t)item * a.n_ct);
    const uint32_t col0 = ct * (VEC * LANES) + lane * VEC;
    const bool active = col0 < a.D;

    if (item >= a.n_chunks) {  // a row without edges: y = delta (or 0)
        if (active) {
            const uint32_t row = __ldg(a.empty_row + (item - a.n_chunks));
            const size_t o = (size_t)row * a.D + col0;
            Acc<VEC> v;
            v.zero();
            if (a.delta) v.load_nc(a.delta + o);
            v.store_cs(out_row(a, row) + col0);
        }
        return;
    }
    // slot(buf, u) = base + ((buf * U + u) * BLOCK + tid) * SLOT : consecutive threads, consecutive slots
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(stage_raw) + threadIdx.x * SLOT;
    const u64* xcol = a.x + col0;

    const uint32_t c = item;
    uint32_t e = c << a.chunk_shift;
    const uint32_t end = min(a.n_edges, e + (1u << a.chunk_shift));
    const uint32_t cn = __ldg(a.chunk_nz + c);
    uint32_t k = cn & ~CGB_END_FLAG;
    bool head_open = (cn & CGB_END_FLAG) != 0;
    bool open = false, have_head = false;
    uint32_t head_row = 0;
    Acc<VEC> acc;
    acc.zero();

    // consume one staged sub-batch: add its rows in edge order, finishing rows where the end flag is set
    auto consume = [&](uint32_t buf, uint32_t fmask, uint32_t cnt) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if ((uint32_t)u < cnt) {
                if (active) acc.add(lds_acc<VEC>(smem_base + (buf * U + u) * (BLOCK * SLOT)));
                open = true;
                if ((fmask >> u) & 1u) {
                    const uint32_t row = __ldg(a.nz_row + k);
                    if (!head_open) {
                        if (active) {
                            const size_t o = (size_t)row * a.D + col0;
                            if (a.delta) {
                                Acc<VEC> d;
                                d.load_nc(a.delta + o);
                                acc.add(d);
                            }
                            acc.store_cs(out_row(a, row) + col0);
                        }
                    } else {
                        if (active) acc.store(a.piece_head + (size_t)c * a.D + col0);
                        head_row = row;
                        have_head = true;
                    }
                    acc.zero();
                    head_open = false;
                    open = false;
                    ++k;
                }
            }
        }
    };

    uint32_t my = (e + lane < end) ? __ldg(a.colf + e + lane) : 0u;
    uint32_t issued = 0;                 // sub-batches issued so far
    uint32_t q_mask[NBUF - 1], q_cnt[NBUF - 1];  // flags / sizes of the sub-batches still in flight (oldest first)
#pragma unroll
    for (int i = 0; i < NBUF - 1; ++i) { q_mask[i] = 0; q_cnt[i] = 0; }
    uint32_t inflight = 0;
    while (e < end) {
        const uint32_t n = min((uint32_t)LANES, end - e);
        const uint32_t nxt = (e + LANES + lane < end) ? __ldg(a.colf + e + LANES + lane) : 0u;  // prefetch indices
        for (uint32_t k0 = 0; k0 < n; k0 += U) {
            const uint32_t buf = issued % NBUF;
            uint32_t fmask = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t id = __shfl_sync(mask, my, k0 + u, LANES);
                if (k0 + u < n) {
                    fmask |= (id >> 31) << u;
                    if (active) {
                        const u64* src = xcol + (size_t)(id & ~CGB_END_FLAG) * a.D;
                        const uint32_t dst = smem_base + (buf * U + u) * (BLOCK * SLOT);
                        if (VEC == 2) cp_async_16(dst, src);
                        else cp_async_8(dst, src);
                    }
                }
            }
            cp_async_commit();
            ++issued;
            if (inflight == NBUF - 1) {  // ring full: retire the oldest sub-batch
                cp_async_wait<NBUF - 1>();
                consume((issued - NBUF) % NBUF, q_mask[0], q_cnt[0]);
#pragma unroll
                for (int i = 0; i + 1 < NBUF - 1; ++i) { q_mask[i] = q_mask[i + 1]; q_cnt[i] = q_cnt[i + 1]; }
                --inflight;
            }
            q_mask[inflight] = fmask;
            q_cnt[inflight] = min((uint32_t)U, n - k0);
            ++inflight;
        }
        e += LANES;
        my = nxt;
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < NBUF - 1; ++i)
        if ((uint32_t)i < inflight) consume((issued - inflight + i) % NBUF, q_mask[i], q_cnt[i]);

    if (have_head) piece_arrive<VEC, LANES>(a, head_row, ct, col0, active, lane, mask);
    if (open) {  // the chunk ends inside a row
        if (active) acc.store((head_open ? a.piece_head : a.piece_tail) + (size_t)c * a.D + col0);
        piece_arrive<VEC, LANES>(a, __ldg(a.nz_row + k), ct, col0, active, lane, mask);
    }
}

__global__ void __launch_bounds__(256) set_end_flags_kernel(const uint32_t* __restrict__ rowptr, uint32_t n_rows,
                                                            uint32_t* __restrict__ colf) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const uint32_t b = rowptr[r], e = rowptr[r + 1];
    if (e > b) colf[e - 1] |= CGB_END_FLAG;
}

// y[j, :] = (idx[j] == NO_ROW ? 0 : x[idx[j], :]) + delta[j, :]
template <int VEC, int LANES>
__global__ void __launch_bounds__(256) expand_rows_kernel(const uint32_t* __restrict__ idx, uint64_t n_out,
                                                          const u64* __restrict__ x, const u64* __restrict__ delta,
                                                          u64* __restrict__ y, uint32_t D, uint32_t n_ct) {
    constexpr int GROUPS = 256 / LANES;
    const int lane = threadIdx.x & (LANES - 1);
    const uint64_t gid = (uint64_t)blockIdx.x * GROUPS + threadIdx.x / LANES;
    if (gid >= n_out * n_ct) return;
    const uint64_t j = gid / n_ct;
    const uint32_t ct = (uint32_t)(gid - j * n_ct);
    const uint32_t col0 = ct * (VEC * LANES) + lane * VEC;
    if (col0 >= D) return;
    const uint32_t src = __ldg(idx + j);
    Acc<VEC> v;
    if (src == CGB_NO_ROW) v.zero();
    else v.load_nc(x + (size_t)src * D + col0);
    if (delta) {
        Acc<VEC> d;
        d.load_nc(delta + j * D + col0);
        v.add(d);
    }
    v.store_cs(y + j * D + col0);
}

// segmented sum over dst-sorted rows; dup: write the sum to every row of the segment
template <int VEC, int LANES, int U>
__global__ void __launch_bounds__(256) segsum_kernel(const uint32_t* __restrict__ segptr, uint32_t n_seg,
                                                     const u64* __restrict__ in, u64* __restrict__ out, uint32_t D,
                                                     uint32_t n_ct, int dup) {
    constexpr int GROUPS = 256 / LANES;
    const int lane = threadIdx.x & (LANES - 1);
    const unsigned mask = group_mask<LANES>(threadIdx.x & 31);
    const uint64_t gid = (uint64_t)blockIdx.x * GROUPS + threadIdx.x / LANES;
    if (gid >= (uint64_t)n_seg * n_ct) return;
    const uint32_t s = (uint32_t)(gid / n_ct);
    const uint32_t ct = (uint32_t)(gid - (uint64_t)s * n_ct);
    const uint32_t col0 = ct * (VEC * LANES) + lane * VEC;
    const bool active = col0 < D;
    const uint32_t b = __ldg(segptr + s), e = __ldg(segptr + s + 1);
    Acc<VEC> acc = accumulate_edges<VEC, LANES, U, true>(nullptr, in, b, e, D, col0, active, lane, mask);
    if (!active) return;
    if (dup) {
        for (uint32_t r = b; r < e; ++r) acc.store_cs(out + (size_t)r * D + col0);
    } else {
        acc.store_cs(out + (size_t)s * D + col0);
    }
}

struct Shape {
    int vec, lanes;
    uint32_t n_ct;
};
inline Shape pick_shape(uint32_t D, bool aligned16) {
    Shape s;
    s.vec = (D % 2 == 0 && aligned16) ? 2 : 1;
    uint32_t need = (D + s.vec - 1) / s.vec;
    int lanes = 2;
    while ((uint32_t)lanes < need && lanes < 32) lanes *= 2;
    s.lanes = lanes;
    s.n_ct = (D + s.vec * lanes - 1) / (s.vec * lanes);
    return s;
}
inline bool is_aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool is_aligned32(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 31) == 0; }
// 256-bit shape for the edge-balanced kernel: D a multiple of 4 columns and every base 32-byte aligned
inline Shape pick_shape4(uint32_t D) {
    Shape s;
    s.vec = 4;
    const uint32_t need = D / 4;
    int lanes = 2;
    while ((uint32_t)lanes < need && lanes < 32) lanes *= 2;
    s.lanes = lanes;
    s.n_ct = (D + 4 * lanes - 1) / (4 * lanes);
    return s;
}
template <typename F>
int dispatch_shape4(const Shape& s, F&& f) {
#define CGB_CASE4(L, U_)                          \
    if (s.vec == 4 && s.lanes == L) {             \
        f(std::integral_constant<int, 4>(), std::integral_constant<int, L>(), std::integral_constant<int, U_>()); \
        return 0;                                 \
    }
    CGB_CASE4(2, 2) CGB_CASE4(4, 4) CGB_CASE4(8, 8) CGB_CASE4(16, 8) CGB_CASE4(32, 8)
#undef CGB_CASE4
    return -1;
}

template <typename F>
int dispatch_shape(const Shape& s, F&& f) {
#define CGB_CASE(V, L, U_)                         \
    if (s.vec == V && s.lanes == L) {              \
        f(std::integral_constant<int, V>(), std::integral_constant<int, L>(), std::integral_constant<int, U_>()); \
        return 0;                                  \
    }
    CGB_CASE(1, 2, 2) CGB_CASE(1, 4, 4) CGB_CASE(1, 8, 8) CGB_CASE(1, 16, 8) CGB_CASE(1, 32, 8)
    CGB_CASE(2, 2, 2) CGB_CASE(2, 4, 4) CGB_CASE(2, 8, 8) CGB_CASE(2, 16, 8) CGB_CASE(2, 32, 8)
#undef CGB_CASE
    return -1;
}

int build_slices(cgb_ctx* ctx, cgb_csr* c, const uint32_t* h_rowptr) {
    std::vector<uint32_t> srow, sbeg, sfirst, scnt, lid;
    uint32_t n_long = 0;
    for (uint32_t v = 0; v < c->n_rows; ++v) {
        uint32_t deg = h_rowptr[v + 1] - h_rowptr[v];
        if (deg <= CGB_LONG_ROW) continue;
        uint32_t ns = (deg + CGB_SLICE_EDGES - 1) / CGB_SLICE_EDGES;
        uint32_t first = (uint32_t)srow.size();
        for (uint32_t k = 0; k < ns; ++k) {
            srow.push_back(v);
            sbeg.push_back(h_rowptr[v] + k * CGB_SLICE_EDGES);
            sfirst.push_back(first);
            scnt.push_back(ns);
            lid.push_back(n_long);
        }
        ++n_long;
    }
    c->n_long_rows = n_long;
    c->n_slices = (uint32_t)srow.size();
    if (c->n_slices == 0) return 0;
    size_t b = (size_t)c->n_slices * sizeof(uint32_t);
    uint32_t** dst[5] = {&c->d_slice_row, &c->d_slice_begin, &c->d_slice_first, &c->d_slice_count, &c->d_long_id};
    const std::vector<uint32_t>* src[5] = {&srow, &sbeg, &sfirst, &scnt, &lid};
    for (int i = 0; i < 5; ++i) {
        CGB_CHECK_CUDA(ctx, cudaMalloc((void**)dst[i], b));
        CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(*dst[i], src[i]->data(), b, cudaMemcpyHostToDevice, ctx->stream));
    }
    CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}


// chunk schedule for gather_chunk_kernel (host side, once per graph)
int build_chunks(cgb_ctx* ctx, cgb_csr* c, const uint32_t* h_rowptr) {
    std::vector<uint32_t> nz, empty;
    nz.reserve(c->n_rows);
    for (uint32_t v = 0; v < c->n_rows; ++v) {
        if (h_rowptr[v + 1] > h_rowptr[v]) nz.push_back(v);
        else empty.push_back(v);
    }
    c->n_nz = (uint32_t)nz.size();
    c->n_empty = (uint32_t)empty.size();
    static const int forced_shift = getenv("CGB_CHUNK_SHIFT") ? atoi(getenv("CGB_CHUNK_SHIFT")) : 0;
    c->chunk_shift = forced_shift >= 4 && forced_shift <= 12 ? (uint32_t)forced_shift
                     : (c->n_edges >= CGB_BIG_GRAPH_EDGES ? CGB_CHUNK_SHIFT_BIG
                        : c->n_edges >= (1ull << 21) ? CGB_CHUNK_SHIFT_SMALL
                        : c->n_edges >= (1ull << 18) ? 5u : 4u);  // small graphs are latency bound: shorter serial chains
    const uint64_t chunk_edges = 1ull << c->chunk_shift;
    c->n_chunks = (uint32_t)((c->n_edges + chunk_edges - 1) / chunk_edges);
    std::vector<uint32_t> chunk_nz(c->n_chunks);
    uint32_t k = 0;
    for (uint32_t ch = 0; ch < c->n_chunks; ++ch) {
        const uint32_t e = ch << c->chunk_shift;
        while (h_rowptr[nz[k] + 1] <= e) ++k;  // advance to the non-empty row that holds edge e
        chunk_nz[ch] = k | (h_rowptr[nz[k]] < e ? CGB_END_FLAG : 0u);
    }
    auto upload = [&](uint32_t** dst, const std::vector<uint32_t>& src) -> int {
        if (src.empty()) return 0;
        CGB_CHECK_CUDA(ctx, cudaMalloc((void**)dst, src.size() * sizeof(uint32_t)));
        CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                            ctx->stream));
        return 0;
    };
    int rc;
    if ((rc = upload(&c->d_nz_row, nz))) return rc;
    c->h_nz_row = nz;
    if ((rc = upload(&c->d_empty_row, empty))) return rc;
    if ((rc = upload(&c->d_chunk_nz, chunk_nz))) return rc;
    if (c->n_edges) {
        CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&c->d_colf, c->n_edges * sizeof(uint32_t)));
        CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(c->d_colf, c->d_col, c->n_edges * sizeof(uint32_t),
                                            cudaMemcpyDeviceToDevice, ctx->stream));
        set_end_flags_kernel<<<(c->n_rows + 255) / 256, 256, 0, ctx->stream>>>(c->d_rowptr, c->n_rows, c->d_colf);
        CGB_CHECK_LAUNCH(ctx, "set_end_flags_kernel");
    }
    CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

bool use_row_schedule() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CGB_GATHER_IMPL");
        v = (e && std::string(e) == "rows") ? 1 : 0;
    }
    return v == 1;
}

}  // namespace

extern "C" {

int cgb_csr_create(cgb_ctx* ctx, const uint32_t* h_rowptr, const uint32_t* h_col, uint32_t n_rows, uint64_t n_edges,
                   uint32_t n_src_rows, cgb_csr** out) {
    CGB_REQUIRE(ctx, out && h_rowptr && (h_col || n_edges == 0), "cgb_csr_create: null argument");
    CGB_REQUIRE(ctx, n_edges < 0xFFFFFFFFull, "cgb_csr_create: edge count must fit 32 bits");
    CGB_REQUIRE(ctx, h_rowptr[0] == 0 && h_rowptr[n_rows] == n_edges, "cgb_csr_create: rowptr does not span the edges");
    for (uint32_t v = 0; v < n_rows; ++v)
        CGB_REQUIRE(ctx, h_rowptr[v] <= h_rowptr[v + 1], "cgb_csr_create: rowptr not monotone");
    for (uint64_t e = 0; e < n_edges; ++e)
        CGB_REQUIRE(ctx, h_col[e] < n_src_rows, "cgb_csr_create: column index out of range");
    CGB_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    cgb_csr* c = new cgb_csr();
    c->n_rows = n_rows;
    c->n_edges = n_edges;
    c->n_src_rows = n_src_rows;
    CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&c->d_rowptr, ((size_t)n_rows + 1) * sizeof(uint32_t)));
    CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&c->d_col, std::max<size_t>(n_edges, 1) * sizeof(uint32_t)));
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(c->d_rowptr, h_rowptr, ((size_t)n_rows + 1) * sizeof(uint32_t),
                                        cudaMemcpyHostToDevice, ctx->stream));
    if (n_edges)
        CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(c->d_col, h_col, n_edges * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                            ctx->stream));
    CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = build_slices(ctx, c, h_rowptr);
    if (rc) return rc;
    rc = build_chunks(ctx, c, h_rowptr);
    if (rc) return rc;
    *out = c;
    return CGB_OK;
}

int cgb_csr_create_device(cgb_ctx* ctx, const uint32_t* d_rowptr, const uint32_t* d_col, uint32_t n_rows,
                          uint64_t n_edges, uint32_t n_src_rows, cgb_csr** out) {
    CGB_REQUIRE(ctx, out && d_rowptr && (d_col || n_edges == 0), "cgb_csr_create_device: null argument");
    CGB_REQUIRE(ctx, n_edges < 0xFFFFFFFFull, "cgb_csr_create_device: edge count must fit 32 bits");
    CGB_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<uint32_t> h_rowptr((size_t)n_rows + 1);
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(h_rowptr.data(), d_rowptr, h_rowptr.size() * sizeof(uint32_t),
                                        cudaMemcpyDeviceToHost, ctx->stream));
    CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CGB_REQUIRE(ctx, h_rowptr[0] == 0 && h_rowptr[n_rows] == n_edges, "cgb_csr_create_device: rowptr does not span the edges");
    cgb_csr* c = new cgb_csr();
    c->n_rows = n_rows;
    c->n_edges = n_edges;
    c->n_src_rows = n_src_rows;
    CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&c->d_rowptr, ((size_t)n_rows + 1) * sizeof(uint32_t)));
    CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&c->d_col, std::max<size_t>(n_edges, 1) * sizeof(uint32_t)));
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(c->d_rowptr, d_rowptr, ((size_t)n_rows + 1) * sizeof(uint32_t),
                                        cudaMemcpyDeviceToDevice, ctx->stream));
    if (n_edges)
        CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(c->d_col, d_col, n_edges * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                            ctx->stream));
    CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = build_slices(ctx, c, h_rowptr.data());
    if (rc) return rc;
    rc = build_chunks(ctx, c, h_rowptr.data());
    if (rc) return rc;
    *out = c;
    return CGB_OK;
}

int cgb_csr_destroy(cgb_ctx* ctx, cgb_csr* c) {
    if (!c) return CGB_OK;
    if (ctx) cudaSetDevice(ctx->device);
    cudaFree(c->d_rowptr); cudaFree(c->d_col);
    cudaFree(c->d_slice_row); cudaFree(c->d_slice_begin); cudaFree(c->d_slice_first);
    cudaFree(c->d_slice_count); cudaFree(c->d_long_id); cudaFree(c->d_counters); cudaFree(c->d_partial);
    cudaFree(c->d_colf); cudaFree(c->d_nz_row); cudaFree(c->d_empty_row); cudaFree(c->d_chunk_nz);
    cudaFree(c->d_chunk_ctr); cudaFree(c->d_piece); cudaFree(c->d_sig_done);
    for (void* p : c->retired) cudaFree(p);
    delete c;
    return CGB_OK;
}
uint64_t cgb_csr_num_edges(const cgb_csr* c) { return c ? c->n_edges : 0; }
uint32_t cgb_csr_num_rows(const cgb_csr* c) { return c ? c->n_rows : 0; }
const uint32_t* cgb_csr_rowptr(const cgb_csr* c) { return c ? c->d_rowptr : nullptr; }
const uint32_t* cgb_csr_col(const cgb_csr* c) { return c ? c->d_col : nullptr; }

static int gather_impl(cgb_ctx* ctx, const cgb_csr* csr_c, const uint64_t* d_x, const uint64_t* d_delta, uint64_t* d_y,
                       uint32_t D, int n_blk, uint64_t* const* blk_base, const uint32_t* blk_off, bool compact = false,
                       const SigArgs* sig = nullptr) {
    cgb_csr* csr = const_cast<cgb_csr*>(csr_c);
    CGB_REQUIRE(ctx, csr && d_x && (d_y || n_blk > 0 || sig) && D > 0, "cgb_gather_sum: null argument");
    CGB_REQUIRE(ctx, (const void*)d_x != (const void*)d_y, "cgb_gather_sum: y must not alias x");
    if (csr->n_rows == 0) return CGB_OK;
    bool al = is_aligned16(d_x) && is_aligned16(d_delta) && is_aligned16(d_y);
    for (int t = 0; t < n_blk; ++t) al = al && is_aligned16(blk_base[t]);
    if (sig)
        for (uint32_t t = 0; t < sig->n_sig; ++t) al = al && is_aligned16(sig->base[t]);
    Shape s = pick_shape(D, al);
    // CGB_GATHER_VEC=4|2: 256-bit accesses when the shape allows (D % 4 == 0, 32-byte aligned bases), else 128 / 64 bit
    static const int want_vec = getenv("CGB_GATHER_VEC") ? atoi(getenv("CGB_GATHER_VEC")) : CGB_GATHER_DEFAULT_VEC;
    static const bool async_impl = getenv("CGB_GATHER_IMPL") && std::string(getenv("CGB_GATHER_IMPL")) == "async";
    bool al32 = is_aligned32(d_x) && is_aligned32(d_delta) && is_aligned32(d_y);
    for (int t = 0; t < n_blk; ++t) al32 = al32 && is_aligned32(blk_base[t]);
    if (sig)
        for (uint32_t t = 0; t < sig->n_sig; ++t) al32 = al32 && is_aligned32(sig->base[t]);
    const bool vec4 = want_vec == 4 && D % 4 == 0 && al32 && !use_row_schedule() && !async_impl;
    if (vec4) s = pick_shape4(D);
    CGB_REQUIRE(ctx, n_blk == 0 || !use_row_schedule(), "cgb_gather_sum_blocks: needs the edge-balanced schedule");
    if (!use_row_schedule()) {
        CGB_REQUIRE(ctx, csr->n_src_rows < CGB_END_FLAG, "cgb_gather_sum: source rows must fit 31 bits");
        if (csr->n_chunks) {
            // grow-only scratch of the CSR handle.  An outgrown buffer is RETIRED, not freed: a captured CUDA graph or a launch
            // still in flight may hold its address; everything is released in cgb_csr_destroy.  (A handle serves one stream at
            // a time -- see the contract in include/cognn_b200.h.)
            size_t need = 2 * (size_t)csr->n_chunks * D;
            if (need > csr->piece_words) {
                if (csr->d_piece) csr->retired.push_back(csr->d_piece);
                csr->d_piece = nullptr;
                csr->piece_words = 0;
                CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&csr->d_piece, need * sizeof(u64)));
                csr->piece_words = need;
            }
            uint32_t need_ctr = csr->n_chunks * s.n_ct;
            if (need_ctr > csr->chunk_ctr_len) {
                if (csr->d_chunk_ctr) csr->retired.push_back(csr->d_chunk_ctr);
                csr->d_chunk_ctr = nullptr;
                csr->chunk_ctr_len = 0;
                CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&csr->d_chunk_ctr, need_ctr * sizeof(uint32_t)));
                CGB_CHECK_CUDA(ctx, cudaMemsetAsync(csr->d_chunk_ctr, 0, need_ctr * sizeof(uint32_t), ctx->stream));
                csr->chunk_ctr_len = need_ctr;
            }
        }
        ChunkArgs a;
        a.colf = csr->d_colf; a.chunk_nz = csr->d_chunk_nz; a.nz_row = csr->d_nz_row; a.rowptr = csr->d_rowptr;
        a.empty_row = csr->d_empty_row;
        a.x = (const u64*)d_x; a.delta = (const u64*)d_delta; a.y = (u64*)d_y;
        a.n_chunks = csr->n_chunks; a.n_empty = csr->n_empty; a.n_edges = (uint32_t)csr->n_edges;
        a.chunk_shift = csr->chunk_shift;
        a.D = D; a.n_ct = s.n_ct;
        a.counters = csr->d_chunk_ctr;
        a.piece_head = (u64*)csr->d_piece;
        a.piece_tail = (u64*)csr->d_piece + (size_t)csr->n_chunks * D;
        a.n_blk = n_blk;
        for (int t = 0; t < n_blk; ++t) {
            a.blk_base[t] = (u64*)blk_base[t];
            a.blk_off[t] = blk_off[t];
        }
        if (n_blk) a.blk_off[n_blk] = blk_off[n_blk];
        a.compact = compact ? 1u : 0u;
        const uint64_t total = ((uint64_t)csr->n_chunks + (sig ? sig->e_cum[sig->n_sig] : (compact ? 0u : csr->n_empty))) * s.n_ct;
        if (total == 0) return CGB_OK;
        // A/B knobs kept from the round-1 experiments (DESIGN.md section 5 has the measurements): CGB_GATHER_IMPL=async -> cp.async
        // staged variant (CGB_GATHER_NBUF=2|3); CGB_GATHER_VEC=2 -> 128-bit kernels; CGB_GATHER_IPL=1 -> one index word per lane
        static const bool use_async = getenv("CGB_GATHER_IMPL") && std::string(getenv("CGB_GATHER_IMPL")) == "async";
        static const int nbuf = getenv("CGB_GATHER_NBUF") ? atoi(getenv("CGB_GATHER_NBUF")) : 2;
        SigArgs g;
        if (sig) {
            g = *sig;
            for (uint32_t t = 0; t < g.n_sig; ++t) g.total[t] *= s.n_ct;  // one store per (row, column tile)
        }
        if (use_async) {
            CGB_REQUIRE(ctx, !compact && !sig, "cgb_gather_sum_compact / _signal: not available with CGB_GATHER_IMPL=async");
            int rc = dispatch_shape(s, [&](auto V, auto L, auto U_) {
                constexpr int BLOCK = 128;
                constexpr int VV = decltype(V)::value, LL = decltype(L)::value, UU = decltype(U_)::value;
                constexpr int GROUPS = BLOCK / LL;
                const unsigned blocks = (unsigned)((total + GROUPS - 1) / GROUPS);
                if (nbuf == 3) {
                    constexpr size_t smem = (size_t)3 * UU * BLOCK * VV * 8;
                    auto kfn = gather_chunk_async_kernel<VV, LL, UU, BLOCK, 3>;
                    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    kfn<<<blocks, BLOCK, smem, ctx->stream>>>(a);
                } else {
                    constexpr size_t smem = (size_t)2 * UU * BLOCK * VV * 8;
                    auto kfn = gather_chunk_async_kernel<VV, LL, UU, BLOCK, 2>;
                    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    kfn<<<blocks, BLOCK, smem, ctx->stream>>>(a);
                }
            });
            CGB_REQUIRE(ctx, rc == 0, "cgb_gather_sum: no kernel for this shape");
            CGB_CHECK_LAUNCH(ctx, "gather_chunk_async_kernel");
            ctx->last_kernel = "gather_chunk_async_kernel";
            return CGB_OK;
        }
        if (vec4) {
            int rc4 = dispatch_shape4(s, [&](auto V, auto L, auto U_) {
                constexpr int BLOCK = 128;
                constexpr int VV = decltype(V)::value, LL = decltype(L)::value;
                constexpr int GROUPS = BLOCK / LL;
                constexpr int UU = decltype(U_)::value;
                constexpr int UH = UU >= 4 ? 4 : UU;
                const unsigned blocks = (unsigned)((total + GROUPS - 1) / GROUPS);
                static const int ipl = getenv("CGB_GATHER_IPL") ? atoi(getenv("CGB_GATHER_IPL")) : 2;
                if (LL <= 4 && ipl == 2) {  // 8 (LANES = 4) or 4 (LANES = 2) edges per batch, 4 row loads in flight per lane
                    constexpr int B2 = LL * 2;
                    constexpr int U4 = B2 >= 4 ? 4 : B2;
                    if (sig) gather_chunk_signal_kernel<VV, LL, U4, BLOCK, 1024, 2><<<blocks, BLOCK, 0, ctx->stream>>>(a, g);
                    else if (compact) gather_chunk_kernel<VV, LL, U4, BLOCK, 1024, 2, true><<<blocks, BLOCK, 0, ctx->stream>>>(a);
                    else gather_chunk_kernel<VV, LL, U4, BLOCK, 1024, 2><<<blocks, BLOCK, 0, ctx->stream>>>(a);
                    ctx->last_kernel = LL == 4 ? "gather_chunk_kernel<VEC=4,LANES=4,U=4,128,1024,IPL=2> (256-bit row loads)"
                                               : "gather_chunk_kernel<VEC=4,LANES=2,U=4,128,1024,IPL=2> (256-bit row loads)";
                } else {
                    if (sig) gather_chunk_signal_kernel<VV, LL, UH, BLOCK, 1024, 1><<<blocks, BLOCK, 0, ctx->stream>>>(a, g);
                    else if (compact) gather_chunk_kernel<VV, LL, UH, BLOCK, 1024, 1, true><<<blocks, BLOCK, 0, ctx->stream>>>(a);
                    else gather_chunk_kernel<VV, LL, UH, BLOCK, 1024><<<blocks, BLOCK, 0, ctx->stream>>>(a);
                    ctx->last_kernel = "gather_chunk_kernel<VEC=4,LANES>=8,U=4,128,1024,IPL=1> (256-bit row loads)";
                }
            });
            CGB_REQUIRE(ctx, rc4 == 0, "cgb_gather_sum: no 256-bit kernel for this shape");
            CGB_CHECK_LAUNCH(ctx, "gather_chunk_kernel");
            return CGB_OK;
        }
        int rc = dispatch_shape(s, [&](auto V, auto L, auto U_) {
            constexpr int BLOCK = 128;
            constexpr int VV = decltype(V)::value, LL = decltype(L)::value;
            constexpr int GROUPS = BLOCK / LL;
            constexpr int UU = decltype(U_)::value;
            constexpr int UH = UU >= 8 ? 4 : UU;
            const unsigned blocks = (unsigned)((total + GROUPS - 1) / GROUPS);
            if (sig) gather_chunk_signal_kernel<VV, LL, UH, BLOCK, 1536, 1><<<blocks, BLOCK, 0, ctx->stream>>>(a, g);
            else if (compact) gather_chunk_kernel<VV, LL, UH, BLOCK, 1536, 1, true><<<blocks, BLOCK, 0, ctx->stream>>>(a);
            else gather_chunk_kernel<VV, LL, UH, BLOCK, 1536><<<blocks, BLOCK, 0, ctx->stream>>>(a);
        });
        CGB_REQUIRE(ctx, rc == 0, "cgb_gather_sum: no kernel for this shape");
        CGB_CHECK_LAUNCH(ctx, "gather_chunk_kernel");
        ctx->last_kernel = "gather_chunk_kernel<VEC<=2,...,128,1536,IPL=1> (128-/64-bit row loads)";
        return CGB_OK;
    }
    CGB_REQUIRE(ctx, !compact && !sig, "cgb_gather_sum_compact / _signal: need the edge-balanced schedule");
    if (csr->n_slices) {
        size_t need = (size_t)csr->n_slices * D;
        if (need > csr->partial_words || !is_aligned16(csr->d_partial)) {
            if (csr->d_partial) csr->retired.push_back(csr->d_partial);
            csr->d_partial = nullptr;
            CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&csr->d_partial, need * sizeof(u64)));
            csr->partial_words = need;
        }
        uint32_t need_ctr = csr->n_long_rows * s.n_ct;
        if (need_ctr > csr->counters_len) {
            if (csr->d_counters) csr->retired.push_back(csr->d_counters);
            csr->d_counters = nullptr;
            CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&csr->d_counters, need_ctr * sizeof(uint32_t)));
            CGB_CHECK_CUDA(ctx, cudaMemsetAsync(csr->d_counters, 0, need_ctr * sizeof(uint32_t), ctx->stream));
            csr->counters_len = need_ctr;
        }
    }
    GatherArgs a;
    a.rowptr = csr->d_rowptr; a.col = csr->d_col;
    a.x = (const u64*)d_x; a.delta = (const u64*)d_delta; a.y = (u64*)d_y;
    a.n_rows = csr->n_rows; a.D = D; a.n_ct = s.n_ct;
    a.n_slices = csr->n_slices;
    a.slice_row = csr->d_slice_row; a.slice_begin = csr->d_slice_begin; a.slice_first = csr->d_slice_first;
    a.slice_count = csr->d_slice_count; a.long_id = csr->d_long_id; a.counters = csr->d_counters;
    a.partial = (u64*)csr->d_partial;
    const uint64_t total = ((uint64_t)csr->n_slices + csr->n_rows) * s.n_ct;
    int rc = dispatch_shape(s, [&](auto V, auto L, auto U_) {
        constexpr int GROUPS = 256 / decltype(L)::value;
        uint64_t blocks = (total + GROUPS - 1) / GROUPS;
        gather_sum_kernel<decltype(V)::value, decltype(L)::value, decltype(U_)::value>
            <<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
    });
    CGB_REQUIRE(ctx, rc == 0, "cgb_gather_sum: no kernel for this shape");
    CGB_CHECK_LAUNCH(ctx, "gather_sum_kernel");
    ctx->last_kernel = "gather_sum_kernel (row schedule)";
    return CGB_OK;
}

int cgb_gather_sum(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* d_x, const uint64_t* d_delta, uint64_t* d_y,
                   uint32_t D) {
    return gather_impl(ctx, csr, d_x, d_delta, d_y, D, 0, nullptr, nullptr);
}

int cgb_gather_sum_blocks(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* d_x, const uint64_t* d_delta, uint32_t D,
                          uint32_t n_blocks, uint64_t* const* d_block_base, const uint32_t* block_row_offsets) {
    CGB_REQUIRE(ctx, csr && d_block_base && block_row_offsets, "cgb_gather_sum_blocks: null argument");
    CGB_REQUIRE(ctx, n_blocks >= 1 && n_blocks <= CGB_MAX_BLOCKS, "cgb_gather_sum_blocks: 1..16 blocks");
    CGB_REQUIRE(ctx, block_row_offsets[0] == 0 && block_row_offsets[n_blocks] == csr->n_rows,
                "cgb_gather_sum_blocks: offsets must span the rows");
    return gather_impl(ctx, csr, d_x, d_delta, nullptr, D, (int)n_blocks, d_block_base, block_row_offsets);
}

int cgb_gather_sum_compact(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* d_x, const uint64_t* d_delta, uint64_t* d_y,
                           uint32_t D) {
    CGB_REQUIRE(ctx, csr, "cgb_gather_sum_compact: null argument");
    if (csr->n_nz == 0) return CGB_OK;
    return gather_impl(ctx, csr, d_x, d_delta, d_y, D, 0, nullptr, nullptr, true);
}
int cgb_gather_sum_signal(cgb_ctx* ctx, const cgb_csr* csr_c, const uint64_t* d_x, uint32_t D, uint32_t n_blocks,
                          const uint32_t* block_row_offsets, uint64_t* const* d_block_base, const uint8_t* block_compact,
                          uint32_t* const* d_block_flag, uint32_t flag_value) {
    cgb_csr* csr = const_cast<cgb_csr*>(csr_c);
    CGB_REQUIRE(ctx, csr && d_x && block_row_offsets && d_block_base && block_compact, "cgb_gather_sum_signal: null argument");
    CGB_REQUIRE(ctx, n_blocks >= 1 && n_blocks <= CGB_MAX_SIG, "cgb_gather_sum_signal: 1..32 blocks");
    CGB_REQUIRE(ctx, block_row_offsets[0] == 0 && block_row_offsets[n_blocks] == csr->n_rows,
                "cgb_gather_sum_signal: offsets must span the rows");
    if (csr->n_rows == 0) return CGB_OK;
    if (!csr->d_sig_done) {
        CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&csr->d_sig_done, CGB_MAX_SIG * sizeof(uint32_t)));
        CGB_CHECK_CUDA(ctx, cudaMemsetAsync(csr->d_sig_done, 0, CGB_MAX_SIG * sizeof(uint32_t), ctx->stream));
    }
    SigArgs g;
    memset(&g, 0, sizeof(g));
    g.n_sig = n_blocks;
    g.done = csr->d_sig_done;
    g.value = flag_value;
    const std::vector<uint32_t>& nz = csr->h_nz_row;
    for (uint32_t b = 0; b <= n_blocks; ++b) {
        CGB_REQUIRE(ctx, b == 0 || block_row_offsets[b] >= block_row_offsets[b - 1], "cgb_gather_sum_signal: offsets must ascend");
        g.off[b] = block_row_offsets[b];
    }
    for (uint32_t b = 0; b < n_blocks; ++b) {
        const uint32_t k_lo = (uint32_t)(std::lower_bound(nz.begin(), nz.end(), g.off[b]) - nz.begin());
        const uint32_t k_hi = (uint32_t)(std::lower_bound(nz.begin(), nz.end(), g.off[b + 1]) - nz.begin());
        g.k0[b] = k_lo;
        if (block_compact[b]) g.compact_mask |= 1u << b;
        g.total[b] = block_compact[b] ? (k_hi - k_lo) : (g.off[b + 1] - g.off[b]);
        g.e_cum[b + 1] = g.e_cum[b] + (block_compact[b] ? 0u : (g.off[b + 1] - g.off[b]) - (k_hi - k_lo));
        g.base[b] = (u64*)d_block_base[b];
        g.flag[b] = d_block_flag ? d_block_flag[b] : nullptr;
        CGB_REQUIRE(ctx, g.base[b] || g.total[b] == 0, "cgb_gather_sum_signal: null block buffer");
    }
    // a block without any row to store completes trivially: raise its flag from the stream
    for (uint32_t b = 0; b < n_blocks; ++b)
        if (g.total[b] == 0 && g.flag[b]) {
            int rc = cgb_flag_signal(ctx, g.flag[b], flag_value);
            if (rc) return rc;
        }
    return gather_impl(ctx, csr, d_x, nullptr, nullptr, D, 0, nullptr, nullptr, false, &g);
}
uint32_t cgb_csr_num_nonempty_rows(const cgb_csr* c) { return c ? c->n_nz : 0; }
const uint32_t* cgb_csr_nonempty_rows(const cgb_csr* c) { return c ? c->d_nz_row : nullptr; }

// Pipelined host entry point: slot s = step & 1 owns a device staging area; the H2D of step i+1 runs on its own stream
// while step i computes and step i-1... copies back (PCIe is full duplex), so host-buffer throughput approaches
// max(H2D, D2H, kernel) per step instead of their sum.
int cgb_host_gather_sum_async(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* h_x, const uint64_t* h_delta,
                              uint64_t* h_y, uint32_t D) {
    CGB_REQUIRE(ctx, csr && h_x && h_y && D > 0, "cgb_host_gather_sum_async: null argument");
    auto& P = ctx->pipe;
    CGB_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!P.ready) {
        CGB_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&P.h2d, cudaStreamNonBlocking));
        CGB_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&P.d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CGB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&P.ev_x[i], cudaEventDisableTiming));
            CGB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&P.ev_y[i], cudaEventDisableTiming));
            CGB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&P.ev_out[i], cudaEventDisableTiming));
        }
        P.ready = true;
    }
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t xb = (size_t)csr->n_src_rows * D * sizeof(u64);
    const size_t yb = (size_t)csr->n_rows * D * sizeof(u64);
    const size_t need = align(xb) + align(yb) + (h_delta ? align(yb) : 0);
    const int s = (int)(P.steps & 1);
    if (need > P.bytes[s]) {
        CGB_CHECK_CUDA(ctx, cudaDeviceSynchronize());
        if (P.buf[s]) cudaFree(P.buf[s]);
        P.buf[s] = nullptr;
        P.bytes[s] = 0;
        CGB_CHECK_CUDA(ctx, cudaMalloc(&P.buf[s], need));
        P.bytes[s] = need;
    }
    char* base = (char*)P.buf[s];
    u64* d_x = (u64*)base;
    u64* d_y = (u64*)(base + align(xb));
    u64* d_delta = h_delta ? (u64*)(base + align(xb) + align(yb)) : nullptr;
    // the slot is free once its previous result has left the device
    if (P.steps >= 2) CGB_CHECK_CUDA(ctx, cudaStreamWaitEvent(P.h2d, P.ev_out[s], 0));
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_x, h_x, xb, cudaMemcpyHostToDevice, P.h2d));
    if (h_delta) CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_delta, h_delta, yb, cudaMemcpyHostToDevice, P.h2d));
    CGB_CHECK_CUDA(ctx, cudaEventRecord(P.ev_x[s], P.h2d));
    CGB_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, P.ev_x[s], 0));
    int rc = cgb_gather_sum(ctx, csr, (const uint64_t*)d_x, (const uint64_t*)d_delta, (uint64_t*)d_y, D);
    if (rc) return rc;
    CGB_CHECK_CUDA(ctx, cudaEventRecord(P.ev_y[s], ctx->stream));
    CGB_CHECK_CUDA(ctx, cudaStreamWaitEvent(P.d2h, P.ev_y[s], 0));
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(h_y, d_y, yb, cudaMemcpyDeviceToHost, P.d2h));
    CGB_CHECK_CUDA(ctx, cudaEventRecord(P.ev_out[s], P.d2h));
    ++P.steps;
    return CGB_OK;
}

int cgb_host_sync(cgb_ctx* ctx) {
    auto& P = ctx->pipe;
    if (P.ready) {
        CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(P.h2d));
        CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(P.d2h));
    } else {
        CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return CGB_OK;
}

int cgb_ipc_export(cgb_ctx* ctx, void* d_ptr, void* out_handle64) {
    CGB_REQUIRE(ctx, d_ptr && out_handle64, "cgb_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CGB_CHECK_CUDA(ctx, cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(out_handle64, &h, sizeof(h));
    return CGB_OK;
}
int cgb_ipc_open(cgb_ctx* ctx, const void* handle64, void** d_peer_out) {
    CGB_REQUIRE(ctx, handle64 && d_peer_out, "cgb_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    CGB_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    CGB_CHECK_CUDA(ctx, cudaIpcOpenMemHandle(d_peer_out, h, cudaIpcMemLazyEnablePeerAccess));
    return CGB_OK;
}
int cgb_ipc_close(cgb_ctx* ctx, void* d_peer) {
    if (!d_peer) return CGB_OK;
    CGB_CHECK_CUDA(ctx, cudaIpcCloseMemHandle(d_peer));
    return CGB_OK;
}

int cgb_expand_rows(cgb_ctx* ctx, const uint32_t* d_idx, uint64_t n_out, const uint64_t* d_x,
                    const uint64_t* d_delta, uint64_t* d_y, uint32_t D) {
    CGB_REQUIRE(ctx, d_idx && d_x && d_y && D > 0, "cgb_expand_rows: null argument");
    CGB_REQUIRE(ctx, (const void*)d_x != (const void*)d_y, "cgb_expand_rows: y must not alias x");
    if (n_out == 0) return CGB_OK;
    const Shape s = pick_shape(D, is_aligned16(d_x) && is_aligned16(d_delta) && is_aligned16(d_y));
    const uint64_t total = n_out * s.n_ct;
    int rc = dispatch_shape(s, [&](auto V, auto L, auto) {
        constexpr int GROUPS = 256 / decltype(L)::value;
        uint64_t blocks = (total + GROUPS - 1) / GROUPS;
        expand_rows_kernel<decltype(V)::value, decltype(L)::value><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
            d_idx, n_out, (const u64*)d_x, (const u64*)d_delta, (u64*)d_y, D, s.n_ct);
    });
    CGB_REQUIRE(ctx, rc == 0, "cgb_expand_rows: no kernel for this shape");
    CGB_CHECK_LAUNCH(ctx, "expand_rows_kernel");
    return CGB_OK;
}

int cgb_segsum(cgb_ctx* ctx, const uint32_t* d_segptr, uint32_t n_seg, uint64_t n_in, const uint64_t* d_in,
               uint64_t* d_out, uint32_t D, int dup) {
    CGB_REQUIRE(ctx, d_segptr && d_out && D > 0 && (d_in || n_in == 0), "cgb_segsum: null argument");
    CGB_REQUIRE(ctx, (const void*)d_in != (const void*)d_out, "cgb_segsum: out must not alias in");
    CGB_REQUIRE(ctx, n_in < 0xFFFFFFFFull, "cgb_segsum: row count must fit 32 bits");
    if (n_seg == 0) return CGB_OK;
    const Shape s = pick_shape(D, is_aligned16(d_in) && is_aligned16(d_out));
    const uint64_t total = (uint64_t)n_seg * s.n_ct;
    int rc = dispatch_shape(s, [&](auto V, auto L, auto U_) {
        constexpr int GROUPS = 256 / decltype(L)::value;
        uint64_t blocks = (total + GROUPS - 1) / GROUPS;
        segsum_kernel<decltype(V)::value, decltype(L)::value, decltype(U_)::value>
            <<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_segptr, n_seg, (const u64*)d_in, (u64*)d_out, D, s.n_ct, dup);
    });
    CGB_REQUIRE(ctx, rc == 0, "cgb_segsum: no kernel for this shape");
    CGB_CHECK_LAUNCH(ctx, "segsum_kernel");
    return CGB_OK;
}

int cgb_host_gather_sum(cgb_ctx* ctx, const cgb_csr* csr, const uint64_t* h_x, const uint64_t* h_delta,
                        uint64_t* h_y, uint32_t D) {
    CGB_REQUIRE(ctx, csr && h_x && h_y && D > 0, "cgb_host_gather_sum: null argument");
    const size_t xb = (size_t)csr->n_src_rows * D * sizeof(u64);
    const size_t yb = (size_t)csr->n_rows * D * sizeof(u64);
    const size_t need = xb + yb + (h_delta ? yb : 0) + 3 * 256;
    int rc = cgb_scratch_reserve(ctx, need);
    if (rc) return rc;
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    char* base = (char*)ctx->scratch;
    u64* d_x = (u64*)base;
    u64* d_y = (u64*)(base + align(xb));
    u64* d_delta = h_delta ? (u64*)(base + align(xb) + align(yb)) : nullptr;
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_x, h_x, xb, cudaMemcpyHostToDevice, ctx->stream));
    if (h_delta) CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_delta, h_delta, yb, cudaMemcpyHostToDevice, ctx->stream));
    rc = cgb_gather_sum(ctx, csr, (const uint64_t*)d_x, (const uint64_t*)d_delta, (uint64_t*)d_y, D);
    if (rc) return rc;
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(h_y, d_y, yb, cudaMemcpyDeviceToHost, ctx->stream));
    CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CGB_OK;
}

}  // extern "C"
