// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM on sm_100a, the suspected bound of the flash-attention
// softmax sweep (profiles/r01_flash_attention.md).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tmem_ld_bench.cu
// Each CTA allocates 512 TMEM columns; `warps` warps (quarter = warp % 4) read their 32 lanes x 128 columns `iters` times
// with .32x32b.x32 / .x64 / .x128 loads, either waiting after every load or once per 128 columns.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ uint32_t ld_cols(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_cols<32>(uint32_t taddr) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s ^= v[i];
    return s;
}
template <>
__device__ __forceinline__ uint32_t ld_cols<64>(uint32_t taddr) {
    uint32_t v[64];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
                   "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]),
                   "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]),
                   "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                 : "r"(taddr));
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 64; i++) s ^= v[i];
    return s;
}

// mode 0: x32 + wait each; 1: 4 x x32 then one wait; 2: 2 x x64 then one wait
template <int MODE>
__global__ void __launch_bounds__(256, 1) tmem_ld_kernel(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) { acc ^= ld_cols<32>(base + 32 * c); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
        } else if (MODE == 1) {
            uint32_t a = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) a ^= ld_cols<32>(base + 32 * c);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= a;
        } else {
            uint32_t a = ld_cols<64>(base) ^ ld_cols<64>(base + 64);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= a;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE>
static void run(const char* name, int warps, int iters) {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 256 * 4);
    tmem_ld_kernel<MODE><<<148, warps * 32>>>(10, cyc, sink);
    tmem_ld_kernel<MODE><<<148, warps * 32>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = (double)warps * 32 * 128 * 4 * iters;          // per CTA
    printf("%-28s warps=%d  %8.1f cycles per 128-col row block per warp  %6.1f B/clk/SM  (%s)\n", name, warps,
           (double)h[0] / iters, bytes / (double)h[0], cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    for (int warps : {1, 4, 8}) {
        run<0>("x32 + wait per load", warps, 2000);
        run<1>("4 x x32, one wait", warps, 2000);
        run<2>("2 x x64, one wait", warps, 2000);
    }
    return 0;
}
