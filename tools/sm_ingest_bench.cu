// How fast can ONE SM pull bytes from HBM, and does it depend on how many SMs pull at the same time? Sizes the decode weight /
// KV streams: if an SM is capped near 1/148 of the HBM bandwidth, every HBM-bound kernel must load all SMs equally.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sm_ingest tools/sm_ingest_bench.cu && /tmp/sm_ingest
// Each CTA streams its own contiguous slice of a 2 GB buffer (never re-read: every byte comes from DRAM) through a per-warp
// cp.async (LDGSTS, 16 B per lane) ring of DEPTH stages of 2 KB, or with cp.async.bulk (TMA engine, 8 KB pieces, mbarrier).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DEPTH>
__global__ void __launch_bounds__(256) ldgsts_stream(const uint8_t* __restrict__ src, size_t per_cta, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint8_t* base = src + (size_t)blockIdx.x * per_cta + (size_t)warp * (per_cta / 8);
    const int n = (int)(per_cta / 8 / 2048);                  // 2 KB chunks per warp
    const uint32_t ring = smem_u32(sm) + warp * DEPTH * 2048;
    auto issue = [&](int c, int s) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + s * 2048 + j * 512 + lane * 16),
                         "l"(base + (size_t)c * 2048 + j * 512 + lane * 16) : "memory");
    };
    for (int s = 0; s < DEPTH; s++) { if (s < n) issue(s, s); asm volatile("cp.async.commit_group;" ::: "memory"); }
    uint32_t acc = 0;
    int st = 0;
    for (int c = 0; c < n; c++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        __syncwarp();
        acc ^= *reinterpret_cast<const uint32_t*>(sm + warp * DEPTH * 2048 + st * 2048 + lane * 16);
        __syncwarp();
        if (c + DEPTH < n) issue(c + DEPTH, st);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (++st == DEPTH) st = 0;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// the skinny-GEMM access pattern: a CTA owns 16 rows of `row_bytes`; warp w takes the 128-byte chunks w, w + 8, ... of every row;
// one warp instruction copies 8 rows x 64 B (lane = 4 * row + 16-byte quad), four instructions per 2 KB stage
template <int DEPTH>
__global__ void __launch_bounds__(256) ldgsts_rows(const uint8_t* __restrict__ src, int row_bytes, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const uint8_t* base = src + (size_t)blockIdx.x * 16 * row_bytes;
    const int n = row_bytes / 128 / 8;                        // chunks per warp
    const uint32_t ring = smem_u32(sm) + warp * DEPTH * 2048;
    auto issue = [&](int c, int s) {
        const size_t k = (size_t)(warp + 8 * c) * 128 + t * 16;
#pragma unroll
        for (int j = 0; j < 4; j++)                           // (row half, k half)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + s * 2048 + j * 512 + lane * 16),
                         "l"(base + (size_t)(g + 8 * (j >> 1)) * row_bytes + k + (j & 1) * 64) : "memory");
    };
    for (int s = 0; s < DEPTH; s++) { if (s < n) issue(s, s); asm volatile("cp.async.commit_group;" ::: "memory"); }
    uint32_t acc = 0;
    int st = 0;
    for (int c = 0; c < n; c++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        __syncwarp();
        acc ^= *reinterpret_cast<const uint32_t*>(sm + warp * DEPTH * 2048 + st * 2048 + lane * 16);
        __syncwarp();
        if (c + DEPTH < n) issue(c + DEPTH, st);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (++st == DEPTH) st = 0;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// one producer thread per CTA: DEPTH slots of PIECE bytes, cp.async.bulk + mbarrier; 4 consumer warps only wait and touch
template <int DEPTH, int PIECE>
__global__ void __launch_bounds__(160) bulk_stream(const uint8_t* __restrict__ src, size_t per_cta, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t full[DEPTH], empty[DEPTH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint8_t* base = src + (size_t)blockIdx.x * per_cta;
    const int n = (int)(per_cta / PIECE);
    if (threadIdx.x == 0) {
        for (int s = 0; s < DEPTH; s++) {
            asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared.b64 [%0], 4;" ::"r"(smem_u32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto wait = [](uint32_t bar, uint32_t ph) {
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(bar), "r"(ph) : "memory");
    };
    if (warp == 4) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int c = 0; c < n; c++) {
                if (c >= DEPTH) wait(smem_u32(&empty[s]), ph ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(PIECE) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sm + s * PIECE)), "l"(base + (size_t)c * PIECE), "r"(PIECE), "r"(smem_u32(&full[s])) : "memory");
                if (++s == DEPTH) { s = 0; ph ^= 1; }
            }
        }
        return;
    }
    uint32_t acc = 0;
    int s = 0; uint32_t ph = 0;
    for (int c = 0; c < n; c++) {
        wait(smem_u32(&full[s]), ph);
        acc ^= *reinterpret_cast<const uint32_t*>(sm + s * PIECE + threadIdx.x * 16);
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        if (++s == DEPTH) { s = 0; ph ^= 1; }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <typename F>
static float time_us(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(0); cudaDeviceSynchronize();                               // warm-up on one region, timed pass on another (nothing in L2)
    cudaEventRecord(a); f(1); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms * 1e3f;
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const size_t total = 3ull << 30;
    uint8_t* buf; uint32_t* sink;
    cudaMalloc(&buf, total); cudaMalloc(&sink, 4); cudaMemset(buf, 1, total);
    const size_t per_cta = 4ull << 20;                        // 4 MB per CTA: launch overhead negligible
    printf("LDGSTS ring (8 warps x DEPTH x 2 KB per CTA), 4 MB per CTA\n");
    auto run_l = [&](auto depth_tag, int ctas, int cta_per_sm_hint) {
        constexpr int D = decltype(depth_tag)::value;
        const int smem = 8 * D * 2048 * (cta_per_sm_hint == 1 ? 1 : 1);
        cudaFuncSetAttribute(ldgsts_stream<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        // force 1 CTA/SM by padding the dynamic smem request when asked
        const int req = cta_per_sm_hint == 1 ? 120 * 1024 : smem;
        float us = time_us([&](int r) { ldgsts_stream<D><<<ctas, 256, req>>>(buf + (size_t)r * ctas * per_cta, per_cta, sink); });
        double gb = (double)ctas * per_cta / us / 1e3;
        printf("  depth %2d (%3d KB in flight/CTA)  CTAs %4d (%d/SM)  %8.1f us  %7.0f GB/s  = %6.1f KB/us per CTA\n", D, D * 16, ctas,
               cta_per_sm_hint, us, gb, gb / ctas * 1e3 / 1e3);
    };
    for (int ctas : {1, 8, 37, 74, 111, 148}) run_l(std::integral_constant<int, 4>{}, ctas, 1);
    for (int ctas : {74, 148, 296}) run_l(std::integral_constant<int, 4>{}, ctas, 2);
    printf("short kernels, one launch each: contiguous 256 KB per CTA vs 16 rows x 16 KB (the down_proj tile), depth 4, 2 CTAs/SM\n");
    for (int ctas : {148, 192, 296, 592}) {
        cudaFuncSetAttribute(ldgsts_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        float us = time_us([&](int r) { ldgsts_stream<4><<<ctas, 256, 8 * 4 * 2048>>>(buf + (size_t)r * ctas * 262144, 262144, sink); });
        float us2 = time_us([&](int r) { ldgsts_rows<4><<<ctas, 256, 8 * 4 * 2048>>>(buf + (size_t)r * ctas * 262144, 16384, sink); });
        printf("  CTAs %4d  contiguous %6.1f us (%5.0f GB/s)   16 rows x 16 KB %6.1f us (%5.0f GB/s)\n", ctas, us, ctas * 262144.0 / us / 1e3, us2,
               ctas * 262144.0 / us2 / 1e3);
    }
    printf("skinny-GEMM pattern (16 rows per CTA, 8 rows x 64 B per instruction), depth 4 = 64 KB in flight per CTA, 2 CTAs/SM\n");
    for (int row_bytes : {6144, 16384, 262144}) {
        for (int ctas : {148, 296, 592, 1184}) {
            cudaFuncSetAttribute(ldgsts_rows<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            float us = time_us([&](int r) { ldgsts_rows<4><<<ctas, 256, 8 * 4 * 2048>>>(buf + (size_t)r * ctas * 16 * row_bytes, row_bytes, sink); });
            double gb = (double)ctas * 16 * row_bytes / us / 1e3;
            printf("  row %6d B  CTAs %4d  %8.1f us  %7.0f GB/s\n", row_bytes, ctas, us, gb);
        }
    }
    return 0;
    printf("cp.async.bulk ring (DEPTH x PIECE per CTA, 1 CTA/SM)\n");
    auto run_b = [&](auto d_tag, auto p_tag, int ctas) {
        constexpr int D = decltype(d_tag)::value, P = decltype(p_tag)::value;
        cudaFuncSetAttribute(bulk_stream<D, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        float us = time_us([&](int r) { bulk_stream<D, P><<<ctas, 160, 120 * 1024 > D * P ? 120 * 1024 : D * P>>>(buf + (size_t)r * ctas * per_cta, per_cta, sink); });
        double gb = (double)ctas * per_cta / us / 1e3;
        printf("  %2d x %5d B (%3d KB in flight)  CTAs %4d  %8.1f us  %7.0f GB/s  = %6.1f KB/us per CTA\n", D, P, D * P / 1024, ctas, us, gb,
               gb / ctas);
    };
    for (int ctas : {1, 8, 37, 148}) run_b(std::integral_constant<int, 8>{}, std::integral_constant<int, 8192>{}, ctas);
    for (int ctas : {1, 37, 148}) run_b(std::integral_constant<int, 16>{}, std::integral_constant<int, 8192>{}, ctas);
    for (int ctas : {1, 37, 148}) run_b(std::integral_constant<int, 6>{}, std::integral_constant<int, 32768>{}, ctas);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
