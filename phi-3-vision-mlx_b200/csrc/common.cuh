// Shared device helpers for the phi3-b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#ifndef __CUDA_ARCH__
#define P3_HOST 1
#endif

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ---- error plumbing (host) -------------------------------------------------------------
extern "C" const char* p3_last_error(void);
void p3_set_error(const char* fmt, ...);
#define P3_CHECK_ARG(cond, ...)                                    \
    do { if (!(cond)) { p3_set_error(__VA_ARGS__); return -1; } } while (0)
#define P3_CHECK_LAUNCH(name)                                                              \
    do { cudaError_t e_ = cudaGetLastError();                                              \
         if (e_ != cudaSuccess) { p3_set_error("%s: %s", name, cudaGetErrorString(e_)); return -2; } } while (0)

#define P3_TRACE_CTAS 1024
unsigned long long* p3_trace_slot();                       // api.cu: next trace slot or null
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void trace_stamp(unsigned long long* tr, int cta, int i) {
    if (tr && threadIdx.x == 0 && cta < P3_TRACE_CTAS) tr[(size_t)cta * 8 + i] = gtimer();
}
__device__ __forceinline__ void trace_meta(unsigned long long* tr, int cta, int kind) {
    if (tr && threadIdx.x == 0 && cta < P3_TRACE_CTAS) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[(size_t)cta * 8 + 5] = smid; tr[(size_t)cta * 8 + 6] = (unsigned long long)kind; tr[(size_t)cta * 8 + 7] = gridDim.x * gridDim.y * gridDim.z;
    }
}

// ---- small device utilities ---------------------------------------------------------------
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    bf162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    bf162 v = *reinterpret_cast<bf162*>(&u);
    return __bfloat1622float2(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global -> shared (LDGSTS). src_bytes==0 zero-fills.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes = 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// streaming variant: L2 evict-first (weights / KV pages are read once per step and must not displace
// the lines other kernels prefetched for their successors)
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void cp_async16_stream(uint32_t dst, const void* src, uint64_t pol) {
    // NOTE: with a [R + UR + imm] shared address (a warp-uniform stage base) ptxas 12.9 places the 64-bit
    // policy descriptor in an odd uniform register (desc[UR1]) and the hardware rejects the LDGSTS as an
    // illegal instruction. Callers keep the stage base in a vector register (stage index + zero*tid);
    // the build greps the SASS for odd descriptors (Makefile `check-sass`).
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
// each CTA of a grid pulls its slice of [ptr, ptr+bytes) into L2 (the next kernel's weights)
__device__ __forceinline__ void l2_prefetch_slice(const uint8_t* ptr, int64_t bytes, int64_t cta, int64_t n_cta, int tid, int nthreads) {
    const int64_t lines = (bytes + 127) / 128, per = (lines + n_cta - 1) / n_cta;
    const int64_t end = (cta + 1) * per < lines ? (cta + 1) * per : lines;
    for (int64_t i = cta * per + tid; i < end; i += nthreads)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + i * 128));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Programmatic dependent launch (PDL): a kernel launched with the attribute may start while its
// stream predecessor drains; everything that reads the predecessor's output must come after
// pdl_wait(). Both are no-ops for a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#include <cstdlib>
inline bool p3_pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("P3_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
// launch with the programmatic-stream-serialization attribute (decode-path kernels)
template <typename... KArgs, typename... Args>
inline cudaError_t p3_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = p3_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: remember it per device, not per process
struct P3DevFlags {
    bool set[64] = {};
    bool& cur() { int d = 0; cudaGetDevice(&d); return set[(d < 0 || d >= 64) ? 0 : d]; }
};

// Paged KV pool addressing. One pool per layer:
//   pool[page][kv(0=K,1=V)][head][slot(0..PAGE-1)][head_dim]   bf16
#define P3_PAGE 64
__device__ __forceinline__ size_t kv_page_elems(int n_heads, int hd) { return (size_t)2 * n_heads * P3_PAGE * hd; }
__device__ __forceinline__ const bf16* kv_tile_ptr(const bf16* pool, int page, int kv, int head, int n_heads, int hd) {
    return pool + (size_t)page * kv_page_elems(n_heads, hd) + ((size_t)kv * n_heads + head) * P3_PAGE * hd;
}
