// Quantised-cache decode attention (head_dim 96): 4-bit g32 prompt pages are dequantised IN REGISTERS,
// straight into mma.sync fragments — no shared-memory round trip, no extra barrier.
// Replaces the reference's "dequantise the whole prompt every step, then concatenate" (phi.py:536-539).
//
// Fragment construction without ldmatrix: the k index of an MMA is a free permutation as long as both
// operands agree, so
//   * QK^T: lane (g,t) takes the 4 consecutive dims 16ks+4t..+3 of its key row -> one 16-bit load of codes
//     (4 nibbles) per (row, k-step); the Q fragments are loaded with the same permutation;
//   * PV  : output column n=g of d-tile dt is mapped to dim 12g+dt, so a lane needs 12 consecutive dims of
//     4 key rows -> three 16-bit loads per row; the accumulators are un-permuted when they are written out.
// Dequantisation is bf16(q*scale + bias) with one rounding: nibble -> bf16 by OR-ing into 0x4300 (128+q),
// HSUB2 128, HFMA2 with the row's (scale, bias) — bit-identical to kv_quantize_kernel / the oracle rule.
// bf16 pages (the partial last prompt page and generated tokens) go through the same fragment layout
// with plain shared-memory loads.
#include "attn_common.cuh"
#include "../../include/phi3_b200.h"

#define Q4_D 96
#define QA_STAGES 6          // phase A: quantised pages, 8 KB stages
#define QA_STAGE 8192
#define QB_STAGES 2          // phase B: bf16 pages + the new tokens, 24 KB stages (same shared memory)

__device__ __forceinline__ uint32_t deq_pair(uint32_t n_lo, uint32_t n_hi, uint32_t s2, uint32_t b2) {
    // (n_lo, n_hi) in 0..15 -> bf16x2 (n*s + b) with a single rounding
    uint32_t x = 0x43004300u | n_lo | (n_hi << 16);                     // (128+n_lo, 128+n_hi), exact in bf16
    bf162 y = __hsub2(*reinterpret_cast<bf162*>(&x), __float2bfloat162_rn(128.f));
    bf162 z = __hfma2(y, *reinterpret_cast<bf162*>(&s2), *reinterpret_cast<bf162*>(&b2));
    return *reinterpret_cast<uint32_t*>(&z);
}
__device__ __forceinline__ uint32_t dup_lo(uint32_t w) { return __byte_perm(w, w, 0x1010); }   // (lo, lo)
__device__ __forceinline__ uint32_t dup_hi(uint32_t w) { return __byte_perm(w, w, 0x3232); }   // (hi, hi)

__global__ void __launch_bounds__(128, 3) attn_decode_q4_kernel(AttnParams p) {
    constexpr int D = Q4_D, CPR = D / 8, STAGE_B = 2 * 64 * D * 2;    // phase-B stage: one bf16 K+V tile
    constexpr int QC = 64 * D / 2, QM = 64 * (D / 32) * 4;            // 3072 B codes, 768 B meta per (page, kv, head)
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sQ = smem;                                                 // [16][96] bf16, linear
    uint8_t* sRing = smem + 16 * D * 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (p.n_heads / p.n_kv);
    pdl_trigger();
    pdl_wait();
    const int past = p.past_dev ? *p.past_dev : p.past_host;
    const int crow = b / p.row_div;
    const int kv0 = p.kv_start ? p.kv_start[crow] : 0;
    const int s_total = past + p.L;
    const int32_t* bt = p.block_table + (size_t)crow * p.bt_stride;

    for (int idx = tid; idx < 16 * CPR; idx += 128) {
        int r = idx / CPR, c = idx % CPR;
        const bf16* src = p.q + ((size_t)b * p.L + (r < p.L ? r : 0)) * p.ldq + h * D + c * 8;
        cp_async16(smem_u32(sQ) + r * (D * 2) + c * 16, src, r < p.L ? 16 : 0);
    }
    const int t_first = kv0 / 64, t_end = (past + 63) / 64;
    const int nt_all = max(t_end - t_first, 0);
    const int n_lo = t_first + (int)(((long long)nt_all * split) / p.n_splits);
    const int n_hi = t_first + (int)(((long long)nt_all * (split + 1)) / p.n_splits);
    const int n_cached = n_hi - n_lo;
    const bool has_present = (split == p.n_splits - 1);
    const int n_iter = n_cached + (has_present ? 1 : 0);
    const uint64_t pol = l2_evict_first_policy();
    const size_t head_elems = (size_t)P3_PAGE * D, page_elems = 2 * (size_t)p.n_kv * head_elems;

    const int vz = p.zero * tid;                                        // 0, but not provably uniform: see AttnParams::zero
    // tiles of this CTA in position order: nA quantised pages, then bf16 pages, then (last split) the new tokens
    const int nA = min(max(p.n_quant / 64 - n_lo, 0), n_cached);
    const int nB = n_iter - nA;
    auto issue_a = [&](int ia) {                                        // [K codes | V codes | K meta | V meta] = 7680 B
        if (ia < nA) {
            const uint32_t st = smem_u32(sRing) + ((ia + vz) % QA_STAGES) * QA_STAGE;
            const int page = bt[n_lo + ia];
            const uint8_t* kc = p.qcodes + ((size_t)page * 2 * p.n_kv + kvh) * QC;
            const uint8_t* vc = kc + (size_t)p.n_kv * QC;
            const uint8_t* km = reinterpret_cast<const uint8_t*>(p.qmeta) + ((size_t)page * 2 * p.n_kv + kvh) * QM;
            const uint8_t* vm = km + (size_t)p.n_kv * QM;
            for (int i = tid; i < QC / 16; i += 128) {
                cp_async16_stream(st + i * 16, kc + i * 16, pol);
                cp_async16_stream(st + QC + i * 16, vc + i * 16, pol);
            }
            if (tid < QM / 16) {
                cp_async16_stream(st + 2 * QC + tid * 16, km + tid * 16, pol);
                cp_async16_stream(st + 2 * QC + QM + tid * 16, vm + tid * 16, pol);
            }
        }
        cp_async_commit();
    };
    auto issue_b = [&](int ib) {                                        // bf16 page (linear [64][96] K then V) or the new tokens
        const int it = nA + ib;
        if (it < n_iter) {
            const uint32_t st = smem_u32(sRing) + ((ib + vz) % QB_STAGES) * STAGE_B;
            if (it < n_cached) {
                const int page = bt[n_lo + it];
                const bf16* kp = p.pool + (size_t)page * page_elems + (size_t)kvh * head_elems;
                const bf16* vp = kp + (size_t)p.n_kv * head_elems;
                for (int i = tid; i < 64 * CPR; i += 128) {
                    cp_async16_stream(st + i * 16, kp + i * 8, pol);
                    cp_async16_stream(st + 64 * D * 2 + i * 16, vp + i * 8, pol);
                }
            } else {                                                    // the L new tokens, rows >= L zero
                for (int idx = tid; idx < 16 * CPR; idx += 128) {
                    int r = idx / CPR, c = idx % CPR;
                    size_t tok = (size_t)b * p.L + (r < p.L ? r : 0);
                    cp_async16(st + r * (D * 2) + c * 16, p.k + tok * p.ldk + kvh * D + c * 8, r < p.L ? 16 : 0);
                    cp_async16(st + 64 * D * 2 + r * (D * 2) + c * 16, p.v + tok * p.ldv + kvh * D + c * 8, r < p.L ? 16 : 0);
                }
            }
        }
        cp_async_commit();
    };
    cp_async_commit();                                                  // group 0: the Q tile
#pragma unroll
    for (int i = 0; i < QA_STAGES - 1; i++) issue_a(i);

    uint32_t qa[D / 16][4];
    float o[D / 8][4];
#pragma unroll
    for (int dt = 0; dt < D / 8; dt++)
#pragma unroll
        for (int j = 0; j < 4; j++) o[dt][j] = 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    // which of the lane's 12 output dims (12g .. 12g+11) fall into the next 32-wide quantisation group
    const int grpA = (12 * g) >> 5, dt_b = 32 * (grpA + 1) - 12 * g;   // dims dt >= dt_b use group grpA+1

    bool q_loaded = false;
    auto load_q = [&]() {
#pragma unroll
        for (int ks = 0; ks < D / 16; ks++) {                           // permuted Q fragments: dims 16ks+4t..+3
            uint2 lo = *reinterpret_cast<const uint2*>(sQ + g * (D * 2) + (ks * 16 + 4 * t) * 2);
            uint2 hi = *reinterpret_cast<const uint2*>(sQ + (g + 8) * (D * 2) + (ks * 16 + 4 * t) * 2);
            qa[ks][0] = lo.x; qa[ks][1] = hi.x; qa[ks][2] = lo.y; qa[ks][3] = hi.y;
        }
        q_loaded = true;
    };
    auto process = [&](const uint8_t* st, int it, bool quant, bool present) {
        if (present && warp != 0) return;
        const int key0 = present ? 0 : warp * 16;                       // this warp's 16 keys inside the tile
        // ---- S = Q K^T
        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        if (quant) {
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                const int r = key0 + nt * 8 + g;
                const uint8_t* row = st + r * (D / 2);
                const uint32_t* meta = reinterpret_cast<const uint32_t*>(st + 2 * QC) + r * 3;
#pragma unroll
                for (int grp = 0; grp < 3; grp++) {
                    const uint32_t mw = meta[grp], s2 = dup_lo(mw), b2 = dup_hi(mw);
#pragma unroll
                    for (int kk = 0; kk < 2; kk++) {
                        const int ks = grp * 2 + kk;
                        const uint32_t hw = *reinterpret_cast<const uint16_t*>(row + ks * 8 + 2 * t);
                        const uint32_t b0 = deq_pair(hw & 15, (hw >> 4) & 15, s2, b2);
                        const uint32_t b1 = deq_pair((hw >> 8) & 15, (hw >> 12) & 15, s2, b2);
                        mma_bf16_16816(s[nt], qa[ks], b0, b1);
                    }
                }
            }
        } else {
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                const uint8_t* row = st + (key0 + nt * 8 + g) * (D * 2);
#pragma unroll
                for (int ks = 0; ks < D / 16; ks++) {
                    uint2 kv = *reinterpret_cast<const uint2*>(row + (ks * 16 + 4 * t) * 2);
                    mma_bf16_16816(s[nt], qa[ks], kv.x, kv.y);
                }
            }
        }
        // ---- mask (boundary tiles only)
        const int j0 = present ? past : (n_lo + it) * 64 + warp * 16;
        if (present || j0 < kv0 || j0 + 16 > past) {
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    int j = j0 + nt * 8 + 2 * t + (e & 1);
                    bool ok;
                    if (present) { int qi = past + ((e & 2) ? g + 8 : g); ok = (j >= kv0) && (j < s_total) && (j <= qi); }
                    else ok = (j >= kv0) && (j < past);
                    if (!ok) s[nt][e] = -INFINITY;
                }
        }
        // ---- online softmax
        float mx0 = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
        float mx1 = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m[0], mx0), mn1 = fmaxf(m[1], mx1);
        const float mu0 = (mn0 == -INFINITY) ? 0.f : mn0 * p.scale_log2, mu1 = (mn1 == -INFINITY) ? 0.f : mn1 * p.scale_log2;
        const float c0 = ex2_approx(m[0] * p.scale_log2 - mu0), c1 = ex2_approx(m[1] * p.scale_log2 - mu1);
        m[0] = mn0; m[1] = mn1;
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            s[nt][0] = ex2_approx(fmaf(s[nt][0], p.scale_log2, -mu0)); s[nt][1] = ex2_approx(fmaf(s[nt][1], p.scale_log2, -mu0));
            s[nt][2] = ex2_approx(fmaf(s[nt][2], p.scale_log2, -mu1)); s[nt][3] = ex2_approx(fmaf(s[nt][3], p.scale_log2, -mu1));
        }
        l[0] = fmaf(l[0], c0, (s[0][0] + s[0][1]) + (s[1][0] + s[1][1]));
        l[1] = fmaf(l[1], c1, (s[0][2] + s[0][3]) + (s[1][2] + s[1][3]));
        if (c0 != 1.f || c1 != 1.f) {
#pragma unroll
            for (int dt = 0; dt < D / 8; dt++) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
        }
        // ---- O += P V ; B fragment column g of d-tile dt holds dim 12g+dt; rows R0..R3 = keys 2t,2t+1,2t+8,2t+9
        const uint32_t a[4] = {pack_bf16(s[0][0], s[0][1]), pack_bf16(s[0][2], s[0][3]),
                               pack_bf16(s[1][0], s[1][1]), pack_bf16(s[1][2], s[1][3])};
        const int R0 = key0 + 2 * t;
        if (quant) {
            const uint8_t* vc = st + QC;
            const uint32_t* vm = reinterpret_cast<const uint32_t*>(st + 2 * QC + QM);
            uint32_t hw[4][3];                                          // 12 nibbles (dims 12g..12g+11) of each of the 4 rows
            uint32_t sA[2], bA[2], sB[2], bB[2];                        // packed (row, row+1) scale/bias, groups grpA / grpA+1
#pragma unroll
            for (int rp = 0; rp < 2; rp++) {
                const int ra = R0 + rp * 8, rb = ra + 1;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    hw[rp * 2][j] = *reinterpret_cast<const uint16_t*>(vc + ra * (D / 2) + 6 * g + 2 * j);
                    hw[rp * 2 + 1][j] = *reinterpret_cast<const uint16_t*>(vc + rb * (D / 2) + 6 * g + 2 * j);
                }
                const uint32_t ma = vm[ra * 3 + grpA], mb = vm[rb * 3 + grpA];
                const uint32_t ma2 = vm[ra * 3 + min(grpA + 1, 2)], mb2 = vm[rb * 3 + min(grpA + 1, 2)];
                sA[rp] = __byte_perm(ma, mb, 0x5410); bA[rp] = __byte_perm(ma, mb, 0x7632);
                sB[rp] = __byte_perm(ma2, mb2, 0x5410); bB[rp] = __byte_perm(ma2, mb2, 0x7632);
            }
#pragma unroll
            for (int dt = 0; dt < D / 8; dt++) {
                const int j = dt >> 2, sh = 4 * (dt & 3);
                const bool useB = dt >= dt_b;
                const uint32_t b0 = deq_pair((hw[0][j] >> sh) & 15, (hw[1][j] >> sh) & 15, useB ? sB[0] : sA[0], useB ? bB[0] : bA[0]);
                const uint32_t b1 = deq_pair((hw[2][j] >> sh) & 15, (hw[3][j] >> sh) & 15, useB ? sB[1] : sA[1], useB ? bB[1] : bA[1]);
                mma_bf16_16816(o[dt], a, b0, b1);
            }
        } else {
            const uint8_t* vt = st + 64 * D * 2;
            uint32_t w[4][6];                                           // dims 12g..12g+11 (6 words) of the 4 rows
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
                const int r = R0 + (rr >> 1) * 8 + (rr & 1);
                const uint2* src = reinterpret_cast<const uint2*>(vt + r * (D * 2) + 24 * g);
                uint2 x0 = src[0], x1 = src[1], x2 = src[2];
                w[rr][0] = x0.x; w[rr][1] = x0.y; w[rr][2] = x1.x; w[rr][3] = x1.y; w[rr][4] = x2.x; w[rr][5] = x2.y;
            }
#pragma unroll
            for (int dt = 0; dt < D / 8; dt++) {
                const uint32_t sel = (dt & 1) ? 0x7632 : 0x5410;
                const uint32_t b0 = __byte_perm(w[0][dt >> 1], w[1][dt >> 1], sel);
                const uint32_t b1 = __byte_perm(w[2][dt >> 1], w[3][dt >> 1], sel);
                mma_bf16_16816(o[dt], a, b0, b1);
            }
        }
    };

    // ---- phase A: quantised pages through the deep 8 KB ring
    for (int ia = 0; ia < nA; ia++) {
        cp_async_wait<QA_STAGES - 2>();
        __syncthreads();
        issue_a(ia + QA_STAGES - 1);
        if (!q_loaded) load_q();
        process(sRing + (ia % QA_STAGES) * QA_STAGE, ia, true, false);
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- phase B: bf16 pages (partial prompt page, generated tokens) and the new tokens
#pragma unroll
    for (int i = 0; i < QB_STAGES - 1; i++) issue_b(i);
    for (int ib = 0; ib < nB; ib++) {
        cp_async_wait<QB_STAGES - 2>();
        __syncthreads();
        issue_b(ib + QB_STAGES - 1);
        if (!q_loaded) load_q();
        process(sRing + (ib % QB_STAGES) * STAGE_B, nA + ib, false, nA + ib >= n_cached);
    }
    cp_async_wait<0>();
    __syncthreads();

    float* sm_o = reinterpret_cast<float*>(sRing);                      // [4][16][D], natural dim order
    float* sm_m = sm_o + 4 * 16 * D;
    float* sm_l = sm_m + 64;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        l[i] += __shfl_xor_sync(0xffffffffu, l[i], 1);
        l[i] += __shfl_xor_sync(0xffffffffu, l[i], 2);
    }
    if (t == 0) {
        sm_m[warp * 16 + g] = m[0] * p.scale_log2; sm_m[warp * 16 + g + 8] = m[1] * p.scale_log2;
        sm_l[warp * 16 + g] = l[0]; sm_l[warp * 16 + g + 8] = l[1];
    }
#pragma unroll
    for (int dt = 0; dt < D / 8; dt++) {                                // accumulator column 2t(+1) of tile dt = dim 24t(+12)+dt
        sm_o[(warp * 16 + g) * D + 24 * t + dt] = o[dt][0];
        sm_o[(warp * 16 + g) * D + 24 * t + 12 + dt] = o[dt][1];
        sm_o[(warp * 16 + g + 8) * D + 24 * t + dt] = o[dt][2];
        sm_o[(warp * 16 + g + 8) * D + 24 * t + 12 + dt] = o[dt][3];
    }
    __syncthreads();
    for (int idx = tid; idx < p.L * D; idx += 128) {
        int r = idx / D, d = idx % D;
        float mm = fmaxf(fmaxf(sm_m[r], sm_m[16 + r]), fmaxf(sm_m[32 + r], sm_m[48 + r]));
        float mu = (mm == -INFINITY) ? 0.f : mm;
        float acc = 0.f, ll = 0.f;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            float f = ex2_approx(sm_m[w * 16 + r] - mu);
            acc += f * sm_o[(w * 16 + r) * D + d];
            ll += f * sm_l[w * 16 + r];
        }
        if (p.n_splits == 1) {
            p.out[((size_t)b * p.L + r) * p.ldo + h * D + d] = __float2bfloat16_rn(ll > 0.f ? acc / ll : 0.f);
        } else {
            size_t base = (((size_t)b * p.n_heads + h) * p.n_splits + split) * 16 + r;
            p.ws_o[base * D + d] = acc;
            if (d == 0) { p.ws_ml[base * 2] = mm; p.ws_ml[base * 2 + 1] = ll; }
        }
    }
    if (p.n_splits > 1 && p.counters) {
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(&p.counters[b * p.n_heads + h], 1) == p.n_splits - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            for (int idx = tid; idx < p.L * D; idx += 128) {
                int r = idx / D, d = idx % D;
                size_t base0 = (((size_t)b * p.n_heads + h) * p.n_splits) * 16 + r;
                float mm = -INFINITY;
                for (int sp = 0; sp < p.n_splits; sp++) mm = fmaxf(mm, __ldcg(&p.ws_ml[(base0 + (size_t)sp * 16) * 2]));
                float mu = (mm == -INFINITY) ? 0.f : mm;
                float acc = 0.f, ll = 0.f;
                for (int sp = 0; sp < p.n_splits; sp++) {
                    size_t bs = base0 + (size_t)sp * 16;
                    float f = ex2_approx(__ldcg(&p.ws_ml[bs * 2]) - mu);
                    acc += f * __ldcg(&p.ws_o[bs * D + d]);
                    ll += f * __ldcg(&p.ws_ml[bs * 2 + 1]);
                }
                p.out[((size_t)b * p.L + r) * p.ldo + h * D + d] = __float2bfloat16_rn(ll > 0.f ? acc / ll : 0.f);
            }
            if (tid == 0) p.counters[b * p.n_heads + h] = 0;
        }
    }
}

int launch_decode_q4_d96(AttnParams& p, cudaStream_t st) {
    dim3 grid(p.n_splits, p.n_heads, p.B);
    const int smem = 16 * Q4_D * 2 + QA_STAGES * QA_STAGE;          // = QB_STAGES * 24576: both phases share it
    static P3DevFlags flags; bool& set = flags.cur();
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(attn_decode_q4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        P3_CHECK_ARG(e == cudaSuccess, "attention_decode_q4: smem attribute: %s", cudaGetErrorString(e));
        set = true;
    }
    p3_launch_pdl(attn_decode_q4_kernel, grid, dim3(128), (size_t)smem, st, p);
    P3_CHECK_LAUNCH("attention_decode_q4");
    return 0;
}
