// exchange.cu -- the mirror-update exchange of one GAS iteration between parties that sit on different GPUs of one box.
//
// Replaces CommSync::sendShareVecVec / recvShareVecVec (include/comm_sync.h:245-277) for the update blocks of
// ss_vertex_centric_algo_kernel.h:835 -> 1067/1090 and the GatherComp additions that consume them
// (optimize-gcn/gcn.h:456-463).  Design:
//
//   * the producer gathers its block for party t in COMPACT form (cgb_gather_sum_compact: one row per destination that
//     has an edge from the producer -- that list is the PosVec both sides hold since preprocessing, ssk.h:507-516), into
//     a staging buffer the consumer has mapped with CUDA IPC;
//   * it then raises a flag in the CONSUMER's memory (cgb_flag_signal: one thread, fence + st.release.sys over NVLink);
//   * the consumer's stream waits for the flag (cgb_flag_wait) and runs cgb_scatter_add_rows, which PULLS the rows
//     straight out of the producer's staging buffer over NVLink and adds them into its vertex rows:
//     v[idx[k], :] += block[k, :].  The block is never copied into local HBM first and never re-read by a separate sum.
//
// The wait is either a stream memory operation (cuStreamWaitValue32: no SM is occupied while waiting) or a one-thread
// kernel with a BOUNDED spin that reports a timeout instead of hanging.  Flags only grow (the value is the step number),
// so a late waiter never misses a signal.
#include <cuda.h>

#include <cstring>

#include "common.cuh"

namespace {

// ---- flags -------------------------------------------------------------------------------------------------------------
__global__ void flag_signal_kernel(uint32_t* flag, uint32_t value) {
    __threadfence_system();  // everything this stream wrote before (own HBM or peer memory) is visible system-wide first
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// bounded spin: ~2 s at 1 us per poll, then the error word is raised and the kernel returns (the caller's result is wrong
// and reported as such, but nothing hangs)
__global__ void flag_wait_kernel(const uint32_t* flag, uint32_t value, uint32_t* err, uint32_t max_polls) {
    uint32_t v = 0;
    for (uint32_t i = 0; i < max_polls; ++i) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - value) >= 0) return;
        __nanosleep(1000);
    }
    if (err) atomicExch(err, 1u);
}

typedef CUresult (*wait_value32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
wait_value32_fn resolve_wait_value32() {
    static wait_value32_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (wait_value32_fn)p;
    }();
    return fn;
}

// ---- pull + add of a compact block ---------------------------------------------------------------------------------------
// Work item = one 16-byte slot of one row (row k, slot s < D / 2), consecutive threads take consecutive slots, so a warp reads
// 512 contiguous bytes of the (possibly remote) block per instruction.  Every thread first issues its U remote loads (items
// i, i + stride, ...), then the U local read-modify-writes, so a CTA keeps U x 16 B x 256 threads = 64 KB in flight against the
// ~2 us NVLink round trip; 32-bit index arithmetic only (a 64-bit division per item made the first version ALU-bound at
// 240 GB/s).  `src` may be peer memory; it is read exactly once, with plain (coherent) loads that bypass L1.
constexpr int SA_THREADS = 256;
constexpr int SA_U = 16;

__device__ __forceinline__ ulonglong2 ld_once_v2(const u64* p) {
    ulonglong2 r;
    asm volatile("ld.global.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}

// D even, 16-byte aligned rows, n * D / 2 < 2^32.  SHIFT >= 0: D / 2 == 1 << SHIFT (no division at all)
template <int SHIFT>
__global__ void __launch_bounds__(SA_THREADS) scatter_add_rows_v2_kernel(const uint32_t* __restrict__ idx, uint32_t total,
                                                                         const u64* src, u64* __restrict__ v, uint32_t D,
                                                                         int assign) {
    const uint32_t slots = D / 2;
    const uint32_t stride = gridDim.x * SA_THREADS;
    for (uint32_t i = blockIdx.x * SA_THREADS + threadIdx.x; i < total; i += SA_U * stride) {
        ulonglong2 r[SA_U];
        uint32_t row[SA_U];
#pragma unroll
        for (int u = 0; u < SA_U; ++u) {
            const uint32_t it = i + (uint32_t)u * stride;
            if (it < total && it >= i) {  // (it >= i: no 32-bit wrap)
                r[u] = ld_once_v2(src + 2 * (size_t)it);  // src is dense n x D: item it <-> words [2 it, 2 it + 2)
                row[u] = __ldg(idx + (SHIFT >= 0 ? (it >> SHIFT) : (it / slots)));
            }
        }
#pragma unroll
        for (int u = 0; u < SA_U; ++u) {
            const uint32_t it = i + (uint32_t)u * stride;
            if (it < total && it >= i) {
                const uint32_t s = SHIFT >= 0 ? (it & ((1u << SHIFT) - 1u)) : (it % slots);
                u64* dst = v + (size_t)row[u] * D + 2 * s;
                ulonglong2 o = assign ? make_ulonglong2(0, 0) : *reinterpret_cast<const ulonglong2*>(dst);
                o.x += r[u].x;
                o.y += r[u].y;
                *reinterpret_cast<ulonglong2*>(dst) = o;
            }
        }
    }
}

// any D / alignment: one u64 per item
__global__ void __launch_bounds__(SA_THREADS) scatter_add_rows_v1_kernel(const uint32_t* __restrict__ idx, uint64_t n,
                                                                         const u64* src, u64* __restrict__ v, uint32_t D,
                                                                         int assign) {
    const uint64_t total = n * D;
    const uint64_t stride = (uint64_t)gridDim.x * SA_THREADS;
    for (uint64_t it = (uint64_t)blockIdx.x * SA_THREADS + threadIdx.x; it < total; it += stride) {
        u64 r;
        asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(src + it));
        u64* dst = v + (size_t)__ldg(idx + it / D) * D + it % D;
        *dst = assign ? r : *dst + r;
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- one protocol round between parties on different GPUs: push into the peer's arena, flag, consume, acknowledge -----------
// Every ordered pair of parties (a "link") owns a double-buffered slot pair in the RECEIVER's arena, a data flag in the
// receiver's memory and an acknowledge flag in the sender's.  The s-th round that carries data over a link uses slot s & 1.
//   push (MODE 0): wait until the peer has consumed round s - 2 (ack >= s - 2), copy the segments of this link into the peer's
//                  slot over NVLink, and the CTA that finishes last raises the peer's data flag to s;
//   recv (MODE 1): wait for the data flag >= s, copy the segments out of the local slot to where the protocol wants them, and
//                  the CTA that finishes last raises the peer's acknowledge flag to s.
// The round number s lives in DEVICE memory (link.seq, advanced by the finishing CTA), never in a kernel argument, so a CUDA
// graph that recorded these launches replays correctly.  All waits are bounded and report through *err.
struct XSeg {
    const u64* src;
    u64* dst;
    uint64_t n_words;
    uint32_t link;
};
struct XLink {
    uint32_t* seq;               // rounds completed on this link in this direction (local)
    uint32_t* done;              // CTA completion counter (local, self-resetting)
    const uint32_t* wait_flag;   // local flag the PEER writes
    uint32_t* signal_flag;       // flag in the peer's memory
    uint32_t n_ctas;             // CTAs of this launch that work on this link
    uint32_t recv;               // 0: push link, 1: recv link
};
struct XArgs {
    XSeg seg[16];
    XLink link[16];
    uint64_t slot_words;
    uint32_t* err;
    uint32_t max_polls;
};

// One launch carries both directions: the push segments come first in blockIdx.y, so their CTAs are dispatched before any
// CTA that waits for a peer's data -- a rank never blocks its own outgoing messages behind its incoming ones.
__global__ void __launch_bounds__(256) peer_round_kernel(const __grid_constant__ XArgs a) {
    const XSeg& sg = a.seg[blockIdx.y];
    const XLink& lk = a.link[sg.link];
    const uint32_t MODE = lk.recv;
    __shared__ uint32_t s_round;
    if (threadIdx.x == 0) {
        uint32_t seq;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seq) : "l"(lk.seq) : "memory");
        const uint32_t s = seq + 1;
        const uint32_t need = MODE == 0 ? s - 2 : s;  // push: the slot's previous round was consumed; recv: this round has arrived
        uint32_t failed = 0;  // a run that already timed out once does not wait again: it ends quickly and reports
        if (a.err) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(failed) : "l"(a.err) : "memory");
        if (!failed && (MODE == 1 || s > 2)) {
            uint32_t v = 0, polls = 0;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(lk.wait_flag) : "memory");
                if ((int32_t)(v - need) >= 0) break;
                if (++polls > a.max_polls) {
                    if (a.err) atomicExch(a.err, MODE == 0 ? 2u : 3u);
                    break;
                }
                __nanosleep(40);
            }
        }
        s_round = s;
    }
    __syncthreads();
    const uint32_t s = s_round;
    const u64* src = sg.src + (MODE == 1 ? (size_t)(s & 1u) * a.slot_words : 0);
    u64* dst = sg.dst + (MODE == 0 ? (size_t)(s & 1u) * a.slot_words : 0);
    const uint64_t n = sg.n_words;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        const uint64_t n2 = n >> 1;
        const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(src);
        ulonglong2* d2 = reinterpret_cast<ulonglong2*>(dst);
        uint64_t k = i;
        for (; k + 3 * stride < n2; k += 4 * stride) {  // four 16-byte loads in flight per thread (NVLink round trip)
            // L2-only loads: in recv mode the slot was written by the peer GPU, and L1 knows nothing about remote writes
            const ulonglong2 v0 = __ldcg(s2 + k), v1 = __ldcg(s2 + k + stride), v2 = __ldcg(s2 + k + 2 * stride),
                             v3 = __ldcg(s2 + k + 3 * stride);
            d2[k] = v0; d2[k + stride] = v1; d2[k + 2 * stride] = v2; d2[k + 3 * stride] = v3;
        }
        for (; k < n2; k += stride) d2[k] = __ldcg(s2 + k);
        if (i == 0 && (n & 1)) dst[n - 1] = __ldcg(src + n - 1);
    } else {
        for (uint64_t k = i; k < n; k += stride) dst[k] = __ldcg(src + k);
    }
    __syncthreads();  // every thread's stores happen-before thread 0's fence below (cumulative): one system fence per CTA
    if (threadIdx.x == 0) {
        __threadfence_system();  // the CTA's stores (peer memory in push mode) are visible system-wide before it reports
        const uint32_t prev = atomicAdd(lk.done, 1u);
        if (prev + 1 == lk.n_ctas) {  // every CTA of this link has finished (and has read seq)
            atomicExch(lk.done, 0u);
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(lk.seq), "r"(s) : "memory");
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(lk.signal_flag), "r"(s) : "memory");
        }
    }
}

}  // namespace

extern "C" {

int cgb_flag_signal(cgb_ctx* ctx, uint32_t* d_flag, uint32_t value) {
    CGB_REQUIRE(ctx, d_flag, "cgb_flag_signal: null flag");
    flag_signal_kernel<<<1, 1, 0, ctx->stream>>>(d_flag, value);
    CGB_CHECK_LAUNCH(ctx, "flag_signal_kernel");
    return CGB_OK;
}

int cgb_flag_wait(cgb_ctx* ctx, const uint32_t* d_flag, uint32_t value, int mode, uint32_t* d_err) {
    CGB_REQUIRE(ctx, d_flag, "cgb_flag_wait: null flag");
    if (mode == 0) {  // stream memory operation: the stream stalls in hardware, no SM is held
        wait_value32_fn fn = resolve_wait_value32();
        CGB_REQUIRE(ctx, fn != nullptr, "cgb_flag_wait: cuStreamWaitValue32 unavailable (use mode 1)");
        CUresult r = fn((CUstream)ctx->stream, (CUdeviceptr)(uintptr_t)d_flag, value, CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) return cgb_fail(ctx, CGB_ERR_CUDA, "cuStreamWaitValue32", std::to_string((int)r).c_str());
        return CGB_OK;
    }
    flag_wait_kernel<<<1, 1, 0, ctx->stream>>>(d_flag, value, d_err, 2000000u);
    CGB_CHECK_LAUNCH(ctx, "flag_wait_kernel");
    return CGB_OK;
}

int cgb_peer_round(cgb_ctx* ctx, const cgb_xseg* segs, uint32_t n_seg, const cgb_xlink* links, uint32_t n_links,
                   uint64_t slot_words, uint32_t ctas_per_seg, uint32_t* d_err) {
    CGB_REQUIRE(ctx, n_seg <= 16 && n_links <= 16 && (n_seg == 0 || (segs && links)), "cgb_peer_round: at most 16 segments / links");
    if (n_seg == 0) return CGB_OK;
    if (ctas_per_seg == 0) ctas_per_seg = 8;
    XArgs a;
    memset(&a, 0, sizeof(a));
    uint32_t per_link[16] = {0};
    bool seen_recv = false;
    for (uint32_t i = 0; i < n_seg; ++i) {
        CGB_REQUIRE(ctx, segs[i].link < n_links && segs[i].src && segs[i].dst, "cgb_peer_round: bad segment");
        const bool r = links[segs[i].link].recv != 0;
        CGB_REQUIRE(ctx, r || !seen_recv, "cgb_peer_round: push segments must precede recv segments");
        seen_recv = seen_recv || r;
        a.seg[i].src = (const u64*)segs[i].src;
        a.seg[i].dst = (u64*)segs[i].dst;
        a.seg[i].n_words = segs[i].n_words;
        a.seg[i].link = segs[i].link;
        per_link[segs[i].link] += ctas_per_seg;
    }
    for (uint32_t l = 0; l < n_links; ++l) {
        CGB_REQUIRE(ctx, per_link[l] > 0, "cgb_peer_round: a link without segments");
        CGB_REQUIRE(ctx, links[l].seq && links[l].done && links[l].wait_flag && links[l].signal_flag, "cgb_peer_round: null link field");
        a.link[l].seq = links[l].seq;
        a.link[l].done = links[l].done;
        a.link[l].wait_flag = links[l].wait_flag;
        a.link[l].signal_flag = links[l].signal_flag;
        a.link[l].n_ctas = per_link[l];
        a.link[l].recv = links[l].recv ? 1u : 0u;
    }
    a.slot_words = slot_words;
    a.err = d_err;
    a.max_polls = 8000000u;  // several seconds: far beyond any skew between ranks inside an iteration (cold first epoch included)
    const dim3 grid(ctas_per_seg, n_seg);
    peer_round_kernel<<<grid, 256, 0, ctx->stream>>>(a);
    CGB_CHECK_LAUNCH(ctx, "peer_round_kernel");
    return CGB_OK;
}

int cgb_scatter_add_rows(cgb_ctx* ctx, const uint32_t* d_idx, uint64_t n, const uint64_t* d_src, uint64_t* d_v, uint32_t D,
                         int assign, uint32_t n_ctas) {
    CGB_REQUIRE(ctx, (d_idx && d_src && d_v) || n == 0, "cgb_scatter_add_rows: null argument");
    CGB_REQUIRE(ctx, D > 0, "cgb_scatter_add_rows: D must be positive");
    if (n == 0) return CGB_OK;
    const bool vec = D % 2 == 0 && aligned16(d_src) && aligned16(d_v) && n * (uint64_t)(D / 2) < 0xFFFFFFFFull;
    const uint64_t items = vec ? n * (D / 2) : n * (uint64_t)D;
    const uint64_t per_cta = (uint64_t)SA_THREADS * (vec ? SA_U : 1);
    uint64_t want = (items + per_cta - 1) / per_cta;
    const uint64_t cap = n_ctas ? n_ctas : (uint64_t)ctx->num_sms * 2;
    if (want > cap) want = cap;
    if (vec) {
        const uint32_t slots = D / 2, total = (uint32_t)items;
        int shift = -1;
        for (int sft = 0; sft < 12; ++sft)
            if (slots == (1u << sft)) shift = sft;
        const u64* src = (const u64*)d_src;
        u64* v = (u64*)d_v;
        switch (shift) {
            case 0: scatter_add_rows_v2_kernel<0><<<(unsigned)want, SA_THREADS, 0, ctx->stream>>>(d_idx, total, src, v, D, assign); break;
            case 1: scatter_add_rows_v2_kernel<1><<<(unsigned)want, SA_THREADS, 0, ctx->stream>>>(d_idx, total, src, v, D, assign); break;
            case 2: scatter_add_rows_v2_kernel<2><<<(unsigned)want, SA_THREADS, 0, ctx->stream>>>(d_idx, total, src, v, D, assign); break;
            case 3: scatter_add_rows_v2_kernel<3><<<(unsigned)want, SA_THREADS, 0, ctx->stream>>>(d_idx, total, src, v, D, assign); break;
            case 4: scatter_add_rows_v2_kernel<4><<<(unsigned)want, SA_THREADS, 0, ctx->stream>>>(d_idx, total, src, v, D, assign); break;
            case 5: scatter_add_rows_v2_kernel<5><<<(unsigned)want, SA_THREADS, 0, ctx->stream>>>(d_idx, total, src, v, D, assign); break;
            case 6: scatter_add_rows_v2_kernel<6><<<(unsigned)want, SA_THREADS, 0, ctx->stream>>>(d_idx, total, src, v, D, assign); break;
            default: scatter_add_rows_v2_kernel<-1><<<(unsigned)want, SA_THREADS, 0, ctx->stream>>>(d_idx, total, src, v, D, assign); break;
        }
    } else
        scatter_add_rows_v1_kernel<<<(unsigned)want, SA_THREADS, 0, ctx->stream>>>(d_idx, n, (const u64*)d_src, (u64*)d_v, D,
                                                                                   assign);
    CGB_CHECK_LAUNCH(ctx, "scatter_add_rows_kernel");
    return CGB_OK;
}

}  // extern "C"
