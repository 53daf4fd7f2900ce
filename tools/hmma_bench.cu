// mma.sync.m16n8k16 bf16 latency / throughput on sm_100a (legacy tensor path), for sizing the decode weight-stream kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hmma_bench tools/hmma_bench.cu && /tmp/hmma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mma(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void k(float* out, long long* cyc, int iters) {
    float acc[CHAINS][4];
    uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
#pragma unroll
    for (int c = 0; c < CHAINS; c++) for (int e = 0; e < 4; e++) acc[c][e] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) mma(acc[c], a, 0x3f803f80u, 0x3f803f80u + c);
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) for (int e = 0; e < 4; e++) s += acc[c][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS>
void run(int warps) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<CHAINS><<<148, warps * 32>>>(out, cyc, iters);
    k<CHAINS><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_warp = (double)h / (iters * CHAINS);
    printf("warps/SM %2d chains %d: %.1f cycles per HMMA per warp, SM rate = 1 HMMA per %.2f cycles (%.0f FMA/clk/SM)\n", warps, CHAINS,
           per_warp, per_warp / warps, 2048.0 * warps / per_warp);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
