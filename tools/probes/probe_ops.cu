// Micro-probes used while designing the kernels (not part of the product): throughput of REDUX.SUM,
// SHFL.BFLY and LDS.128 per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_ops probe_ops.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) probe(int iters, uint32_t *out)
{
    __shared__ uint4 sm[1024];
    uint32_t v[8];
    for (int k = 0; k < 8; k++) v[k] = threadIdx.x * 2654435761u + k;
    if (OP == 2) for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (OP == 0) v[k] = __reduce_add_sync(0xFFFFFFFFu, v[k]) + k;
            if (OP == 1) v[k] += __shfl_xor_sync(0xFFFFFFFFu, v[k], 1 + (k & 3));
            if (OP == 2) { uint4 t = sm[(v[k] + threadIdx.x) & 1023]; v[k] += t.x ^ t.y ^ t.z ^ t.w; }
            if (OP == 3) v[k] = __reduce_min_sync(0xFFFFFFFFu, v[k]) + k;
        }
    }
    uint32_t t = 0;
    for (int k = 0; k < 8; k++) t ^= v[k];
    if (t == 0x12345u) out[0] = t;
}

template <int OP>
void run(const char *name, uint32_t *out)
{
    const int iters = 4096, ctas = 148 * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<OP><<<ctas, 256>>>(16, out);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<OP><<<ctas, 256>>>(iters, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warp_instr = (double)ctas * 8 * iters * 8;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-10s %.3f ms  %.2f warp-instr/clk/SM (at %d MHz nominal)  err=%s\n", name, ms, warp_instr / cyc / 148.0, clk / 1000,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    uint32_t *out; cudaMalloc(&out, 4096);
    run<0>("redux.add", out);
    run<3>("redux.min", out);
    run<1>("shfl.bfly", out);
    run<2>("lds.128", out);
    return 0;
}
