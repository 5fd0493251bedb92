// capi.cu -- context, memory and error plumbing of the C ABI (include/cognn_b200.h).
//
// The reference reports errors with printf + exit(-1) (ss_vertex_centric_algo_kernel.h:794-797, 869-872) and has
// no device code; here every entry point returns a status and the C++ shim in cognn_b200/host reproduces the
// reference behaviour on top.  There is deliberately no CPU fallback: without a CUDA device context creation fails.
#include <mutex>

#include "common.cuh"

static thread_local std::string g_last_error;

int cgb_scratch_reserve(cgb_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return CGB_OK;
    // retired, not freed: launches in flight and captured CUDA graphs may still hold the old address (freed with the context)
    if (ctx->scratch) ctx->retired.push_back(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&ctx->scratch, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&ctx->scratch, want);
    }
    if (e != cudaSuccess) return cgb_fail(ctx, CGB_ERR_NOMEM, "cgb_scratch_reserve", cudaGetErrorString(e));
    ctx->scratch_bytes = want;
    return CGB_OK;
}

extern "C" {

int cgb_ctx_set_matmul_impl(cgb_ctx* ctx, int impl) {
    CGB_REQUIRE(ctx, impl >= -1 && impl <= 2, "cgb_ctx_set_matmul_impl: -1 (environment / auto), 0 auto, 1 imad, 2 tc");
    ctx->matmul_impl = impl;
    return CGB_OK;
}
const char* cgb_ctx_last_kernel(cgb_ctx* ctx) { return ctx ? ctx->last_kernel : ""; }

const char* cgb_version(void) { return "cognn_b200 0.1 (sm_100a)"; }

int cgb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int cgb_ctx_create_on_stream(int device, void* cuda_stream, cgb_ctx** out) {
    if (!out) return CGB_ERR_INVALID;
    *out = nullptr;
    int n = cgb_device_count();
    if (n <= 0 || device < 0 || device >= n) {
        g_last_error = "cgb_ctx_create: no CUDA device (this library has no CPU fallback)";
        return CGB_ERR_NO_DEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        g_last_error = "cgb_ctx_create: cudaSetDevice failed";
        return CGB_ERR_CUDA;
    }
    cgb_ctx* c = new cgb_ctx();
    c->device = device;
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    *out = c;
    return CGB_OK;
}

int cgb_ctx_create(int device, cgb_ctx** out) {
    int rc = cgb_ctx_create_on_stream(device, nullptr, out);
    if (rc) return rc;
    cudaStream_t s;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
        delete *out;
        *out = nullptr;
        g_last_error = "cgb_ctx_create: cudaStreamCreate failed";
        return CGB_ERR_CUDA;
    }
    (*out)->stream = s;
    (*out)->own_stream = true;
    return CGB_OK;
}

int cgb_ctx_destroy(cgb_ctx* ctx) {
    if (!ctx) return CGB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->tc_planes) cudaFree(ctx->tc_planes);
    if (ctx->tc_aux) {
        cudaStreamSynchronize(ctx->tc_aux);
        cudaStreamDestroy(ctx->tc_aux);
        for (int i = 0; i < 9; ++i) cudaEventDestroy(ctx->tc_ev[i]);
    }
    for (void* p : ctx->retired) cudaFree(p);
    if (ctx->pipe.ready) {
        cudaStreamSynchronize(ctx->pipe.h2d);
        cudaStreamSynchronize(ctx->pipe.d2h);
        for (int i = 0; i < 2; ++i) {
            if (ctx->pipe.buf[i]) cudaFree(ctx->pipe.buf[i]);
            cudaEventDestroy(ctx->pipe.ev_x[i]);
            cudaEventDestroy(ctx->pipe.ev_y[i]);
            cudaEventDestroy(ctx->pipe.ev_out[i]);
        }
        cudaStreamDestroy(ctx->pipe.h2d);
        cudaStreamDestroy(ctx->pipe.d2h);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return CGB_OK;
}

int cgb_ctx_sync(cgb_ctx* ctx) {
    CGB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CGB_OK;
}
void* cgb_ctx_stream(cgb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
const char* cgb_last_error(cgb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }
uint64_t cgb_ctx_launch_count(cgb_ctx* ctx) { return ctx ? ctx->launches : 0; }
int cgb_ctx_set_prg_stream_bias(cgb_ctx* ctx, const uint64_t* d_bias) {
    CGB_REQUIRE(ctx, ctx != nullptr, "cgb_ctx_set_prg_stream_bias: null context");
    ctx->prg_bias = d_bias;
    return CGB_OK;
}

int cgb_malloc(cgb_ctx* ctx, size_t bytes, void** d_out) {
    CGB_REQUIRE(ctx, d_out, "cgb_malloc: null argument");
    CGB_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(d_out, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cgb_fail(ctx, CGB_ERR_NOMEM, "cgb_malloc", cudaGetErrorString(e));
    }
    return CGB_OK;
}
int cgb_free(cgb_ctx* ctx, void* d_ptr) {
    if (!d_ptr) return CGB_OK;
    CGB_CHECK_CUDA(ctx, cudaFree(d_ptr));
    return CGB_OK;
}
int cgb_host_alloc(size_t bytes, void** h_out) {
    if (!h_out) return CGB_ERR_INVALID;
    cudaError_t e = cudaMallocHost(h_out, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        g_last_error = std::string("cgb_host_alloc: ") + cudaGetErrorString(e);
        return CGB_ERR_NOMEM;
    }
    return CGB_OK;
}
int cgb_host_free(void* h_ptr) {
    if (h_ptr) cudaFreeHost(h_ptr);
    return CGB_OK;
}
int cgb_memset(cgb_ctx* ctx, void* d_ptr, int value, size_t bytes) {
    CGB_CHECK_CUDA(ctx, cudaMemsetAsync(d_ptr, value, bytes, ctx->stream));
    return CGB_OK;
}
int cgb_h2d(cgb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return CGB_OK;
}
int cgb_d2h(cgb_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return CGB_OK;
}
int cgb_d2d(cgb_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
    CGB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return CGB_OK;
}

}  // extern "C"
