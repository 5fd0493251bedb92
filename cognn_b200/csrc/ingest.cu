// ingest.cu -- graph ingest + index-vector construction on the device.
//
// Reference being served (file:line in /root/reference):
//   graphTilesFromEdgeList / partition file      include/graph_io_util.h:40-208  (vertex -> party map, remote in-degrees 170-175)
//   GraphTile finalize (edge order, degrees)     include/graph.h:607-641
//   index vectors of SSEdgeCentricAlgoKernel     include/ss_vertex_centric_algo_kernel.h:295-534 (with -r 1, ssk.h:412-418)
// The reference walks hash maps on the host (std::unordered_map per vertex and per edge); at the 100M-edge sweep size that is
// the slowest step before the hot path.  Here the whole derivation is a handful of HBM-bound passes:
//   1. per party t: exclusive scan of [tid == t]  -> local index of every vertex (ascending vid inside a party, ssk.h:462-464)
//   2. one pass over the edge list: in-degrees (atomics), keys (destination row << 32 | source row) of the party's out-edges
//      appended through a warp-aggregated cursor, per-row edge counts, border flags
//   3. radix sort of the keys (cub::DeviceRadixSort over exactly the significant bits): rows grouped by destination
//      ascending, sources ascending inside a row -- the (src,dst)-sorted edge order of graph.h:636-641
//   4. exclusive scan of the row counts -> rowptr; low key halves -> col; dummy rule -> in_deg
// The result feeds cgb_csr_create_device directly; only the small per-vertex arrays go back to the host.
// CUB (shipped with the CUDA toolkit) provides the scan and the sort; the passes around them are hand written.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <vector>

#include "common.cuh"

struct cgb_party_graph {
    int T = 0, me = 0;
    uint32_t n_local = 0, n_rows = 0;
    uint64_t n_out_edges = 0;
    std::vector<uint32_t> offsets;  // T + 1 (host): first output row of each destination party
    uint64_t* d_vids = nullptr;
    uint64_t* d_in_deg_raw = nullptr;
    uint64_t* d_in_deg = nullptr;
    uint8_t* d_is_border = nullptr;
    uint32_t* d_rowptr = nullptr;
    uint32_t* d_col = nullptr;
};

namespace {

constexpr int IG_THREADS = 256;

struct IsParty {
    const int64_t* tid;
    int64_t t;
    __host__ __device__ uint32_t operator()(uint64_t v) const { return tid[v] == t ? 1u : 0u; }
};
typedef thrust::transform_iterator<IsParty, thrust::counting_iterator<uint64_t>> PartyFlagIter;

__global__ void __launch_bounds__(IG_THREADS) check_tid_kernel(const int64_t* __restrict__ tid, uint64_t n, int T, int* err) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
        if (tid[v] < 0 || tid[v] >= T) *err = 1;
}

// local_index[v] = scan[v] for the vertices of party t; the last vertex also leaves the party's size
__global__ void __launch_bounds__(IG_THREADS) take_local_index_kernel(const int64_t* __restrict__ tid, const uint32_t* __restrict__ scan,
                                                                     uint64_t n, int t, uint32_t* __restrict__ local_index,
                                                                     uint8_t* __restrict__ tid8, uint32_t* __restrict__ count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const bool mine = tid[v] == t;
        if (mine) {
            local_index[v] = scan[v];
            tid8[v] = (uint8_t)t;
        }
        if (v == n - 1) count[t] = scan[v] + (mine ? 1u : 0u);
    }
}

__global__ void __launch_bounds__(IG_THREADS) vids_kernel(const uint8_t* __restrict__ tid8, const uint32_t* __restrict__ local_index,
                                                         uint64_t n, int me, uint64_t* __restrict__ vids) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
        if (tid8[v] == me) vids[local_index[v]] = v;
}

struct Offsets { uint32_t o[CGB_MAX_BLOCKS + 1]; };

// one pass over the edge list (16-byte loads of the (src, dst) pairs; tid8 / local_index gathers hit L2 for the hubs)
__global__ void __launch_bounds__(IG_THREADS) edge_pass_kernel(const longlong2* __restrict__ edges, uint64_t n_edges, uint64_t n_vertices,
                                                              const uint8_t* __restrict__ tid8, const uint32_t* __restrict__ local_index,
                                                              int me, const Offsets off, uint32_t* __restrict__ in_cnt,
                                                              uint32_t* __restrict__ local_in, uint8_t* __restrict__ is_border,
                                                              uint32_t* __restrict__ row_cnt, unsigned long long* __restrict__ keys,
                                                              unsigned long long* __restrict__ cursor, int* err) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t n_round = (n_edges + 31) / 32 * 32;  // whole warps iterate together (ballot below)
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_round; e += stride) {
        bool emit = false;
        unsigned long long key = 0;
        if (e < n_edges) {
            const longlong2 sd = edges[e];
            if (sd.x < 0 || sd.y < 0 || (uint64_t)sd.x >= n_vertices || (uint64_t)sd.y >= n_vertices) {
                *err = 2;
            } else {
                const int ts = tid8[sd.x], td = tid8[sd.y];
                const uint32_t ld = local_index[sd.y];
                if (td == me) atomicAdd(in_cnt + ld, 1u);  // graph.h:627-632 (local) and graph_io_util.h:170-175 (remote)
                if (ts == me) {
                    const uint32_t ls = local_index[sd.x];
                    const uint32_t row = off.o[td] + ld;
                    key = ((unsigned long long)row << 32) | ls;
                    emit = true;
                    atomicAdd(row_cnt + row, 1u);
                    if (td == me) atomicAdd(local_in + ld, 1u);
                    else is_border[ls] = 1;  // isLocalVertexBorder (graph_io_util.h:169)
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, emit);
        if (m) {
            unsigned long long base = 0;
            const int leader = __ffs(m) - 1;
            if ((int)lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (emit) keys[base + __popc(m & ((1u << lane) - 1))] = key;
        }
    }
}

__global__ void __launch_bounds__(IG_THREADS) col_kernel(const unsigned long long* __restrict__ keys, uint64_t m, uint32_t* __restrict__ col) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) col[i] = (uint32_t)keys[i];
}

// ssk.h:412-418: a local vertex without a LOCAL in-edge gets a dummy self edge and its degrees are incremented; the edge
// carries no value (dropped at Gather via isGatherDstVertexDummy), so only the increment survives
__global__ void __launch_bounds__(IG_THREADS) degrees_kernel(const uint32_t* __restrict__ in_cnt, const uint32_t* __restrict__ local_in,
                                                            uint32_t n_local, uint64_t* __restrict__ in_deg_raw, uint64_t* __restrict__ in_deg) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_local) {
        in_deg_raw[i] = in_cnt[i];
        in_deg[i] = (uint64_t)in_cnt[i] + (local_in[i] == 0 ? 1u : 0u);
    }
}

inline unsigned ig_blocks(const cgb_ctx* ctx, uint64_t n) {
    uint64_t b = (n + IG_THREADS - 1) / IG_THREADS;
    const uint64_t cap = (uint64_t)ctx->num_sms * 8;
    if (b > cap) b = cap;
    return (unsigned)(b ? b : 1);
}

struct Tmp {  // frees whatever is still held when the builder leaves, on success or on error
    std::vector<void*> p;
    ~Tmp() {
        for (void* q : p) cudaFree(q);
    }
    template <typename T>
    cudaError_t alloc(T** out, size_t n) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) p.push_back(q);
        *out = (T*)q;
        return e;
    }
};

}  // namespace

extern "C" {

int cgb_party_graph_destroy(cgb_ctx* ctx, cgb_party_graph* g) {
    if (!g) return CGB_OK;
    if (ctx) cudaSetDevice(ctx->device);
    cudaFree(g->d_vids); cudaFree(g->d_in_deg_raw); cudaFree(g->d_in_deg); cudaFree(g->d_is_border);
    cudaFree(g->d_rowptr); cudaFree(g->d_col);
    delete g;
    return CGB_OK;
}

int cgb_party_graph_build(cgb_ctx* ctx, const int64_t* d_edges, uint64_t n_edges, const int64_t* d_tid, uint64_t n_vertices,
                          int T, int me, cgb_party_graph** out) {
    CGB_REQUIRE(ctx, out && (d_edges || n_edges == 0) && d_tid, "cgb_party_graph_build: null argument");
    CGB_REQUIRE(ctx, T >= 1 && T <= CGB_MAX_BLOCKS && me >= 0 && me < T, "cgb_party_graph_build: 1 <= T <= 16, 0 <= me < T");
    CGB_REQUIRE(ctx, n_vertices >= 1 && n_vertices < 0xFFFFFFFFull, "cgb_party_graph_build: vertex count must fit 32 bits");
    CGB_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_edges) & 15) == 0, "cgb_party_graph_build: edge list must be 16-byte aligned");
    CGB_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Tmp tmp;
    int* d_err = nullptr;
    uint32_t *d_scan = nullptr, *d_local_index = nullptr, *d_count = nullptr;
    uint8_t* d_tid8 = nullptr;
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_err, 1));
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_scan, n_vertices));
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_local_index, n_vertices));
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_tid8, n_vertices));
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_count, (size_t)T));
    CGB_CHECK_CUDA(ctx, cudaMemsetAsync(d_err, 0, sizeof(int), st));
    check_tid_kernel<<<ig_blocks(ctx, n_vertices), IG_THREADS, 0, st>>>(d_tid, n_vertices, T, d_err);
    CGB_CHECK_LAUNCH(ctx, "check_tid_kernel");
    // 1. local index of every vertex inside its party
    size_t scan_bytes = 0;
    {
        PartyFlagIter it(thrust::counting_iterator<uint64_t>(0), IsParty{d_tid, 0});
        CGB_CHECK_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, it, d_scan, n_vertices, st));
    }
    void* d_scan_tmp = nullptr;
    CGB_CHECK_CUDA(ctx, tmp.alloc((uint8_t**)&d_scan_tmp, scan_bytes));
    for (int t = 0; t < T; ++t) {
        PartyFlagIter it(thrust::counting_iterator<uint64_t>(0), IsParty{d_tid, (int64_t)t});
        CGB_CHECK_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_scan_tmp, scan_bytes, it, d_scan, n_vertices, st));
        ctx->launches++;
        take_local_index_kernel<<<ig_blocks(ctx, n_vertices), IG_THREADS, 0, st>>>(d_tid, d_scan, n_vertices, t, d_local_index,
                                                                                  d_tid8, d_count);
        CGB_CHECK_LAUNCH(ctx, "take_local_index_kernel");
    }
    std::vector<uint32_t> count(T);
    int h_err = 0;
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(count.data(), d_count, T * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    CGB_REQUIRE(ctx, h_err == 0, "cgb_party_graph_build: tile id out of range");

    cgb_party_graph* g = new cgb_party_graph();
    struct Guard {  // the half-built result is released unless the build completes
        cgb_ctx* c;
        cgb_party_graph* g;
        ~Guard() { if (g) cgb_party_graph_destroy(c, g); }
    } guard{ctx, g};
    g->T = T;
    g->me = me;
    g->offsets.assign(T + 1, 0);
    for (int t = 0; t < T; ++t) g->offsets[t + 1] = g->offsets[t] + count[t];
    g->n_local = count[me];
    g->n_rows = g->offsets[T];
    const uint32_t n_local = g->n_local, n_rows = g->n_rows;
    CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&g->d_vids, std::max<size_t>(n_local, 1) * sizeof(uint64_t)));
    CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&g->d_in_deg_raw, std::max<size_t>(n_local, 1) * sizeof(uint64_t)));
    CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&g->d_in_deg, std::max<size_t>(n_local, 1) * sizeof(uint64_t)));
    CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&g->d_is_border, std::max<size_t>(n_local, 1)));
    CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&g->d_rowptr, ((size_t)n_rows + 1) * sizeof(uint32_t)));
    vids_kernel<<<ig_blocks(ctx, n_vertices), IG_THREADS, 0, st>>>(d_tid8, d_local_index, n_vertices, me, g->d_vids);
    CGB_CHECK_LAUNCH(ctx, "vids_kernel");

    // 2. the edge pass
    uint32_t *d_in_cnt = nullptr, *d_local_in = nullptr, *d_row_cnt = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys_alt = nullptr, *d_cursor = nullptr;
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_in_cnt, n_local));
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_local_in, n_local));
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_row_cnt, (size_t)n_rows + 1));
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_keys, n_edges));
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_cursor, 1));
    CGB_CHECK_CUDA(ctx, cudaMemsetAsync(d_in_cnt, 0, std::max<size_t>(n_local, 1) * sizeof(uint32_t), st));
    CGB_CHECK_CUDA(ctx, cudaMemsetAsync(d_local_in, 0, std::max<size_t>(n_local, 1) * sizeof(uint32_t), st));
    CGB_CHECK_CUDA(ctx, cudaMemsetAsync(d_row_cnt, 0, ((size_t)n_rows + 1) * sizeof(uint32_t), st));
    CGB_CHECK_CUDA(ctx, cudaMemsetAsync(g->d_is_border, 0, std::max<size_t>(n_local, 1), st));
    CGB_CHECK_CUDA(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), st));
    Offsets off;
    for (int t = 0; t <= T; ++t) off.o[t] = g->offsets[t];
    if (n_edges) {
        edge_pass_kernel<<<ig_blocks(ctx, n_edges), IG_THREADS, 0, st>>>((const longlong2*)d_edges, n_edges, n_vertices, d_tid8,
                                                                        d_local_index, me, off, d_in_cnt, d_local_in,
                                                                        g->d_is_border, d_row_cnt, d_keys, d_cursor, d_err);
        CGB_CHECK_LAUNCH(ctx, "edge_pass_kernel");
    }
    unsigned long long m = 0;
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(&m, d_cursor, sizeof(m), cudaMemcpyDeviceToHost, st));
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    CGB_REQUIRE(ctx, h_err == 0, "cgb_party_graph_build: vertex id out of range");
    CGB_REQUIRE(ctx, m < 0xFFFFFFFFull, "cgb_party_graph_build: a party's out-edge count must fit 32 bits");
    g->n_out_edges = m;
    CGB_CHECK_CUDA(ctx, cudaMalloc((void**)&g->d_col, std::max<size_t>(m, 1) * sizeof(uint32_t)));

    // 3. sort by (destination row, source row); only the significant key bits are sorted
    if (m) {
        int row_bits = 1;
        while (row_bits < 32 && (1ull << row_bits) < (unsigned long long)n_rows) ++row_bits;
        CGB_CHECK_CUDA(ctx, tmp.alloc(&d_keys_alt, (size_t)m));
        cub::DoubleBuffer<unsigned long long> db(d_keys, d_keys_alt);
        size_t sort_bytes = 0;
        CGB_CHECK_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, db, (int64_t)m, 0, 32 + row_bits, st));
        void* d_sort_tmp = nullptr;
        CGB_CHECK_CUDA(ctx, tmp.alloc((uint8_t**)&d_sort_tmp, sort_bytes));
        CGB_CHECK_CUDA(ctx, cub::DeviceRadixSort::SortKeys(d_sort_tmp, sort_bytes, db, (int64_t)m, 0, 32 + row_bits, st));
        ctx->launches++;
        col_kernel<<<ig_blocks(ctx, m), IG_THREADS, 0, st>>>(db.Current(), m, g->d_col);
        CGB_CHECK_LAUNCH(ctx, "col_kernel");
    }
    // 4. rowptr and degrees
    {
        size_t b = 0;
        CGB_CHECK_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, b, d_row_cnt, g->d_rowptr, (size_t)n_rows + 1, st));
        void* t2 = nullptr;
        CGB_CHECK_CUDA(ctx, tmp.alloc((uint8_t**)&t2, b));
        CGB_CHECK_CUDA(ctx, cub::DeviceScan::ExclusiveSum(t2, b, d_row_cnt, g->d_rowptr, (size_t)n_rows + 1, st));
        ctx->launches++;
    }
    if (n_local) {
        degrees_kernel<<<(n_local + IG_THREADS - 1) / IG_THREADS, IG_THREADS, 0, st>>>(d_in_cnt, d_local_in, n_local, g->d_in_deg_raw,
                                                                                      g->d_in_deg);
        CGB_CHECK_LAUNCH(ctx, "degrees_kernel");
    }
    CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    guard.g = nullptr;
    *out = g;
    return CGB_OK;
}

int cgb_party_graph_build_host(cgb_ctx* ctx, const int64_t* h_edges, uint64_t n_edges, const int64_t* h_tid, uint64_t n_vertices,
                               int T, int me, cgb_party_graph** out) {
    CGB_REQUIRE(ctx, out && (h_edges || n_edges == 0) && h_tid, "cgb_party_graph_build_host: null argument");
    CGB_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    Tmp tmp;
    int64_t *d_edges = nullptr, *d_tid = nullptr;
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_edges, (size_t)n_edges * 2));
    CGB_CHECK_CUDA(ctx, tmp.alloc(&d_tid, (size_t)n_vertices));
    if (n_edges)
        CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_edges, h_edges, (size_t)n_edges * 16, cudaMemcpyHostToDevice, ctx->stream));
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_tid, h_tid, (size_t)n_vertices * 8, cudaMemcpyHostToDevice, ctx->stream));
    return cgb_party_graph_build(ctx, d_edges, n_edges, d_tid, n_vertices, T, me, out);
}

uint32_t cgb_party_graph_num_local(const cgb_party_graph* g) { return g ? g->n_local : 0; }
uint32_t cgb_party_graph_num_rows(const cgb_party_graph* g) { return g ? g->n_rows : 0; }
uint64_t cgb_party_graph_num_out_edges(const cgb_party_graph* g) { return g ? g->n_out_edges : 0; }
const uint32_t* cgb_party_graph_offsets(const cgb_party_graph* g) { return g ? g->offsets.data() : nullptr; }
const uint64_t* cgb_party_graph_vids(const cgb_party_graph* g) { return g ? g->d_vids : nullptr; }
const uint64_t* cgb_party_graph_in_deg_raw(const cgb_party_graph* g) { return g ? g->d_in_deg_raw : nullptr; }
const uint64_t* cgb_party_graph_in_deg(const cgb_party_graph* g) { return g ? g->d_in_deg : nullptr; }
const uint8_t* cgb_party_graph_is_border(const cgb_party_graph* g) { return g ? g->d_is_border : nullptr; }
const uint32_t* cgb_party_graph_rowptr(const cgb_party_graph* g) { return g ? g->d_rowptr : nullptr; }
const uint32_t* cgb_party_graph_col(const cgb_party_graph* g) { return g ? g->d_col : nullptr; }

int cgb_party_graph_csr(cgb_ctx* ctx, const cgb_party_graph* g, cgb_csr** out) {
    CGB_REQUIRE(ctx, g && out, "cgb_party_graph_csr: null argument");
    return cgb_csr_create_device(ctx, g->d_rowptr, g->d_col, g->n_rows, g->n_out_edges, g->n_local, out);
}

}  // extern "C"
