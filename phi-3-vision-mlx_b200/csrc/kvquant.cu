// 4-bit group-32 affine quantisation of the prompt part of the paged KV cache.
// Replaces mx.quantize(keys.reshape(B*N,-1), group_size=32) of phi.py:532 (bits=4 default).
// Groups are 32 consecutive head-dim elements of one (token, head): 3 per 96-wide head, never
// straddling tokens. Rule (stated in DESIGN.md, mirrors oracle.quantize_q4g32
// in 'b200' mode): the larger-magnitude edge of the group is exact, 0 stays representable,
// scale and bias are rounded to bf16 BEFORE the codes are computed, q = clamp(rint((w-b)/s),0,15).
// Full 64-token pages get 4-bit codes + bf16 (scale,bias); every quantised position (including
// the tail of a partial page) is also written back to the bf16 pool as bf16(q*s+b), so a page
// that is only partly prompt keeps reference semantics while staying in the bf16 pool.
#include "common.cuh"
#include "../../include/phi3_b200.h"

__global__ void kv_quantize_kernel(bf16* __restrict__ pool, uint8_t* __restrict__ qcodes, bf16* __restrict__ qmeta,
                                   const int32_t* __restrict__ block_table, int bt_stride, int n_seq, int n_tokens,
                                   int n_full, int n_kv, int hd) {
    const int gph = hd / 32;                                     // groups per head row
    int64_t total = (int64_t)n_seq * n_tokens * 2 * n_kv * gph;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int gi = (int)(idx % gph);
    int head = (int)((idx / gph) % n_kv);
    int kv = (int)((idx / ((int64_t)gph * n_kv)) % 2);
    int tok = (int)((idx / ((int64_t)gph * n_kv * 2)) % n_tokens);
    int seq = (int)(idx / ((int64_t)gph * n_kv * 2 * n_tokens));
    int page = block_table[(size_t)seq * bt_stride + tok / P3_PAGE], slot = tok % P3_PAGE;
    size_t row = ((size_t)page * 2 + kv) * n_kv + head;         // (page, kv, head)
    bf16* src = pool + (row * P3_PAGE + slot) * hd + gi * 32;
    float w[32];
    uint4 raw[4];
#pragma unroll
    for (int i = 0; i < 4; i++) raw[i] = reinterpret_cast<const uint4*>(src)[i];
    const uint32_t* ru = reinterpret_cast<const uint32_t*>(raw);
    float wmax = -INFINITY, wmin = INFINITY;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float2 f = unpack_bf16(ru[i]);
        w[2 * i] = f.x; w[2 * i + 1] = f.y;
        wmax = fmaxf(wmax, fmaxf(f.x, f.y)); wmin = fminf(wmin, fminf(f.x, f.y));
    }
    bool mask = fabsf(wmin) > fabsf(wmax);
    float scale = fmaxf(__fdiv_rn(__fsub_rn(wmax, wmin), 15.0f), 1e-7f);
    scale = mask ? scale : -scale;
    float edge = mask ? wmin : wmax;
    float q0 = rintf(__fdiv_rn(edge, scale));
    if (q0 != 0.f) scale = __fdiv_rn(edge, q0);
    float bias = (q0 == 0.f) ? 0.f : edge;
    scale = bf16_round(scale); bias = bf16_round(bias);
    if (scale == 0.f) scale = 1e-7f;
    uint32_t packed[4] = {0, 0, 0, 0};
    uint4 deq[4];
    uint32_t* du = reinterpret_cast<uint32_t*>(deq);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float qa = fminf(fmaxf(rintf(__fdiv_rn(__fsub_rn(w[2 * i], bias), scale)), 0.f), 15.f);
        float qb = fminf(fmaxf(rintf(__fdiv_rn(__fsub_rn(w[2 * i + 1], bias), scale)), 0.f), 15.f);
        packed[i / 4] |= ((uint32_t)qa | ((uint32_t)qb << 4)) << (8 * (i % 4));
        du[i] = pack_bf16(__fadd_rn(__fmul_rn(qa, scale), bias), __fadd_rn(__fmul_rn(qb, scale), bias));
    }
#pragma unroll
    for (int i = 0; i < 4; i++) reinterpret_cast<uint4*>(src)[i] = deq[i];
    if (tok < n_full) {
        uint8_t* cdst = qcodes + (row * P3_PAGE + slot) * (hd / 2) + gi * 16;
        *reinterpret_cast<uint4*>(cdst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        bf16* mdst = qmeta + ((row * P3_PAGE + slot) * gph + gi) * 2;
        *reinterpret_cast<uint32_t*>(mdst) = pack_bf16(scale, bias);
    }
}

extern "C" int p3_kv_quantize_q4g32(const void* pool, void* qcodes, void* qmeta, const int32_t* block_table,
                                    int bt_stride, int n_seq, int n_tokens, int n_kv, int hd, cudaStream_t st) {
    P3_CHECK_ARG(hd % 32 == 0, "kv_quantize: head_dim must be a multiple of 32");
    int64_t total = (int64_t)n_seq * n_tokens * 2 * n_kv * (hd / 32);
    if (total == 0) return 0;
    int n_full = (n_tokens / P3_PAGE) * P3_PAGE;
    kv_quantize_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>((bf16*)const_cast<void*>(pool), (uint8_t*)qcodes,
                                                                        (bf16*)qmeta, block_table, bt_stride, n_seq,
                                                                        n_tokens, n_full, n_kv, hd);
    P3_CHECK_LAUNCH("kv_quantize");
    return 0;
}
