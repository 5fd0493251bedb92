// imad_peak.cu -- measures the integer-pipe ceiling for u64 multiply-add (low 64 bits) on this GPU: the roofline
// denominator of matmul_u64 (MEASURED_PEAKS.json only has HBM and bf16).  Pure register work, no memory traffic:
// every thread runs ITER x (8 x 4) independent acc += a * b with the 8 a-operands perturbed each iteration so the
// products cannot be hoisted.  Prints one JSON line.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;

template <int TM, int TN>
__global__ void __launch_bounds__(256) imad_kernel(u64* out, int iters, u64 seed) {
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
        for (int i = 0; i < TM; ++i) fa[i] += (u64)it;
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) s ^= acc[i][j];
    out[t] = s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    u64* out;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(u64));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        imad_kernel<8, 4><<<blocks, threads>>>(out, iters, 0x1234567ull + rep);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        double macs = (double)blocks * threads * iters * 32;
        double rate = macs / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaError_t e = cudaGetLastError();
    printf("{\"kernel\": \"imad_peak_u64_mac\", \"u64_mac_per_s\": %.6e, \"sms\": %d, \"clock_khz\": %d, \"err\": \"%s\"}\n",
           best, prop.multiProcessorCount, prop.clockRate, cudaGetErrorString(e));
    return 0;
}
