// Shared definitions of the attention kernels (attention.cu, attention_q4.cu).
#pragma once
#include "common.cuh"

template <int D> struct Swz;
template <> struct Swz<96> { static __device__ __forceinline__ int f(int c, int r) { return c ^ ((r >> 1) & 3); } };
template <> struct Swz<64> { static __device__ __forceinline__ int f(int c, int r) { return c ^ (r & 7); } };

// byte offset of 16B chunk c of row r inside a [rows][D] bf16 tile
template <int D>
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * (D * 2) + Swz<D>::f(c, r) * 16); }

struct AttnParams {
    const bf16 *q, *k, *v;
    int64_t ldq, ldk, ldv;
    bf16* out; int64_t ldo;
    int B, L, n_heads, n_kv, hd;
    float scale_log2;
    int causal, past;
    int past_host; const int32_t* past_dev;   // decode: past read from device memory when past_dev != NULL (CUDA-graph replay)
    const int32_t* kv_start;
    const bf16* pool;
    const int32_t* block_table; int bt_stride;
    int row_div;
    // decode only
    int n_splits, tiles_per_split;
    float* ws_o; float* ws_ml;
    int zero;                   // always 0 at run time; `zero * tid` keeps stage addresses in vector registers (see cp_async16_stream)
    const uint8_t* l2_prefetch; int64_t l2_prefetch_bytes;   // next kernel's weights (o_proj) pulled into L2 while KV streams
    int* counters;              // [B*n_heads] split-arrival counters (zero on entry, reset by the merging CTA)
    // quantised-cache decode: positions [0, n_quant) (multiple of 64) live in the q4 pools
    int n_quant;
    const uint8_t* qcodes;      // [page][2][n_kv][64][D/2]
    const bf16* qmeta;          // [page][2][n_kv][64][D/32][2] (scale, bias)
    unsigned long long* trace;  // tools/chain_trace.py (null in production)
    int early_fill;             // decode: first KV tiles requested before the dependency wait (P3_ATTN_EARLY=0 turns it off)
};


__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
