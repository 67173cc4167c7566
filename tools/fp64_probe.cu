// fp64_probe.cu -- measures what the FP64 pipe of this GPU actually delivers (throughput of
// independent DFMA / DMUL+DADD streams, latency of a dependent chain, shared-memory fp64 load
// bandwidth).  The tile engine is bound by this pipe, not by HBM, so these are the denominators its
// utilisation is judged against.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool FMA>
__global__ void k_thru(double *out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (FMA) v[i] = __fma_rn(v[i], a, b);
            else v[i] = __dadd_rn(__dmul_rn(v[i], a), b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lat(double *out, long long *cyc, int iters, double a, double b)
{
    double v = threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) v = __fma_rn(v, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_lds(double *out, int iters)
{
    __shared__ double s[8 * 256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) s[i] = i;
    __syncthreads();
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += s[j * 256 + (threadIdx.x + it) % 256];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
float time_ms(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    for (int threads : {256, 1024}) {
        int blocks = sms * (2048 / threads);
        float ms = time_ms([&] { k_thru<8, true><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double ops = (double)blocks * threads * iters * 8;
        printf(", \"dfma_per_s_t%d\": %.4e", threads, ops / (ms * 1e-3));
        ms = time_ms([&] { k_thru<8, false><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf(", \"dmul_dadd_pairs_per_s_t%d\": %.4e", threads, ops / (ms * 1e-3));
    }
    {   // the tile kernel's shape: 1 CTA of 256 threads per SM, ILP 8
        float ms = time_ms([&] { k_thru<8, false><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
        printf(", \"dmul_dadd_pairs_per_s_1cta256\": %.4e", (double)sms * 256 * iters * 8 / (ms * 1e-3));
        ms = time_ms([&] { k_thru<1, false><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
        printf(", \"dmul_dadd_pairs_per_s_1cta256_ilp1\": %.4e", (double)sms * 256 * iters / (ms * 1e-3));
    }
    k_lat<<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf(", \"dfma_dependent_latency_cycles\": %.2f", (double)c / iters);
    {
        float ms = time_ms([&] { k_lds<<<sms * 2, 1024>>>(out, 2048); });
        double bytes = (double)sms * 2 * 1024 * 2048 * 8 * 8;
        printf(", \"lds64_bytes_per_s\": %.4e", bytes / (ms * 1e-3));
    }
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf(", \"sm_clock_khz_attr\": %d}\n", clk);
    return 0;
}
