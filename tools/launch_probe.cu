// Developer probe (under gpurun): where the launch + sync floor of the host-facing step goes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/launch_probe tools/launch_probe.cu
// Measures wall clock per iteration of: launch + cudaStreamSynchronize for an empty kernel, the same with 3.5 KB of
// kernel parameters, with the step's host I/O (49 KB read from / 217 KB written to mapped page-locked memory by 128 CTAs),
// with a ~30 us busy kernel, and the same kernels completed by polling a flag the LAST CTA writes to mapped memory
// instead of cudaStreamSynchronize.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <immintrin.h>

struct Big { char pad[3500]; };

__device__ __forceinline__ void signal_done(unsigned* counter, volatile unsigned* flag, unsigned seq) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned old = atomicAdd(counter, 1u);
        if (old == gridDim.x - 1) {
            *counter = 0;
            __threadfence_system();
            *flag = seq;
        }
    }
}

__global__ void k_empty() {}
__global__ void k_big(const __grid_constant__ Big b, int* sink) { if (sink && b.pad[0] == 77) *sink = 1; }
__global__ void k_io(const float* in, float* out, int n_in, int n_out, long long spin, unsigned* counter, volatile unsigned* flag, unsigned seq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    float acc = 0;
    for (int i = t; i < n_in; i += nt) acc += in[i];
    if (spin > 0) { const long long t0 = clock64(); while (clock64() - t0 < spin) acc += 1e-9f; }
    for (int i = t; i < n_out; i += nt) out[i] = acc + (float)i;
    if (flag) signal_done(counter, flag, seq);
}

template <class F>
static double timeit(int iters, F f) {
    for (int i = 0; i < 200; i++) f(i);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) f(200 + i);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::micro>(t1 - t0).count() / iters;
}

int main() {
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    const int n_in = 4096 * 3, n_out = 4096 * 11 + 4096 * 2 + 1024;
    float *h_in, *h_out, *d_in, *d_out, *g_in, *g_out;
    unsigned *h_flag, *d_flag, *counter;
    cudaHostAlloc(&h_in, n_in * 4, cudaHostAllocMapped); cudaHostAlloc(&h_out, n_out * 4, cudaHostAllocMapped);
    cudaHostAlloc(&h_flag, 64, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&d_in, h_in, 0); cudaHostGetDevicePointer(&d_out, h_out, 0); cudaHostGetDevicePointer(&d_flag, h_flag, 0);
    cudaMalloc(&g_in, n_in * 4); cudaMalloc(&g_out, n_out * 4); cudaMalloc(&counter, 4); cudaMemset(counter, 0, 4);
    memset(h_in, 0, n_in * 4); cudaMemset(g_in, 0, n_in * 4); *h_flag = 0;
    Big big; memset(&big, 0, sizeof big);
    const int N = 3000;
    const long long spin30 = (long long)(30e-6 * 1.9e9);
    printf("empty kernel, launch + cudaStreamSynchronize          %7.2f us\n", timeit(N, [&](int) { k_empty<<<1, 32, 0, st>>>(); cudaStreamSynchronize(st); }));
    printf("3.5 KB parameters, launch + sync                      %7.2f us\n", timeit(N, [&](int) { k_big<<<1, 32, 0, st>>>(big, nullptr); cudaStreamSynchronize(st); }));
    printf("128x128, device I/O, launch + sync                    %7.2f us\n", timeit(N, [&](int) { k_io<<<128, 128, 0, st>>>(g_in, g_out, n_in, n_out, 0, counter, nullptr, 0); cudaStreamSynchronize(st); }));
    printf("128x128, mapped host I/O (49 KB in, 217 KB out), sync %7.2f us\n", timeit(N, [&](int) { k_io<<<128, 128, 0, st>>>(d_in, d_out, n_in, n_out, 0, counter, nullptr, 0); cudaStreamSynchronize(st); }));
    printf("  + 30 us of work, sync                               %7.2f us\n", timeit(N, [&](int) { k_io<<<128, 128, 0, st>>>(d_in, d_out, n_in, n_out, spin30, counter, nullptr, 0); cudaStreamSynchronize(st); }));
    auto poll = [&](unsigned seq) {
        volatile unsigned* f = h_flag;
        long spins = 0;
        while (*f != seq) { _mm_pause(); if (++spins > 200000000L) { printf("flag timeout\n"); break; } }
    };
    printf("128x128, mapped host I/O, completion by FLAG          %7.2f us\n", timeit(N, [&](int i) { k_io<<<128, 128, 0, st>>>(d_in, d_out, n_in, n_out, 0, counter, d_flag, (unsigned)i + 1); poll((unsigned)i + 1); }));
    printf("  + 30 us of work, completion by FLAG                 %7.2f us\n", timeit(N, [&](int i) { k_io<<<128, 128, 0, st>>>(d_in, d_out, n_in, n_out, spin30, counter, d_flag, (unsigned)i + 100001); poll((unsigned)i + 100001); }));
    printf("device I/O, completion by FLAG                        %7.2f us\n", timeit(N, [&](int i) { k_io<<<128, 128, 0, st>>>(g_in, g_out, n_in, n_out, 0, counter, d_flag, (unsigned)i + 200001); poll((unsigned)i + 200001); }));
    cudaStreamSynchronize(st);
    // launch call alone (how long cudaLaunchKernel keeps the CPU)
    {
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 1000; i++) k_big<<<1, 32, 0, st>>>(big, nullptr);
        auto t1 = std::chrono::steady_clock::now();
        cudaStreamSynchronize(st);
        printf("cudaLaunchKernel call alone (queue not full)          %7.2f us\n", std::chrono::duration<double, std::micro>(t1 - t0).count() / 1000);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
