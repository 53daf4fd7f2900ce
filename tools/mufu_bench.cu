// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition on sm_100a (is the flash-attention softmax sweep MUFU-bound?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MIX>   // 0: ex2 only; 1: FFMA + ex2 + FADD per element (the sweep's mix)
__global__ void mufu_kernel(int iters, long long* cycles, float* sink, float sl, float mu) {
    float v[32], acc = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = (float)(threadIdx.x + i) * 1e-3f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 32; i++) {
            float x = MIX ? fmaf(v[i], sl, -mu) : v[i];
            float y;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
            if (MIX) acc += y;
            v[i] = y * 0.5f;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = acc;
#pragma unroll
    for (int i = 0; i < 32; i++) s += v[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MIX>
static void run(const char* name, int warps) {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
    const int iters = 2000;
    mufu_kernel<MIX><<<148, warps * 32>>>(10, cyc, sink, 0.1f, 0.2f);
    mufu_kernel<MIX><<<148, warps * 32>>>(iters, cyc, sink, 0.1f, 0.2f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double per_warp_instr = (double)h[0] / (iters * 32.0);
    printf("%-22s warps/SM=%2d  %6.2f cycles per warp-wide ex2 (per warp)  ->  %5.1f ex2 lanes/clk/SM\n", name, warps, per_warp_instr,
           warps * 32.0 / per_warp_instr);
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    for (int w : {1, 4, 8, 16}) run<0>("ex2 only", w);
    for (int w : {4, 8, 16}) run<1>("ffma + ex2 + fadd", w);
    return 0;
}
