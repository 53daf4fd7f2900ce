// Quantised-cache decode attention (head_dim 96): 4-bit g32 prompt pages go through mma.sync AS CODES — the
// affine map is applied to group sums, not to elements — with no shared-memory round trip and no extra barrier.
// Replaces the reference's "dequantise the whole prompt every step, then concatenate" (phi.py:536-539).
//
//   value(key, d) = s[key][grp(d)] * code + b[key][grp(d)]          (kv_quantize_kernel / mx.quantize, 32-wide groups)
//   QK^T : score = sum_grp  s_grp * (q_grp . code_grp)  +  b_grp * sum(q_grp)
//          -> one accumulator per group (3 x 2 k-steps), codes enter as the exact bf16 numbers 128+code
//             (nibble OR-ed into 0x4300: PRMT + LOP3 per register), the -128 goes into the per-key constant
//             (b - 128 s) that multiplies sum(q_grp); scales are applied to the 16x8 accumulators in fp32.
//   PV   : out[d] = sum_key (P * s[key][grp]) * code  +  sum_key P * b[key][grp]
//          -> the A operand is bf16(P * s_grp), one set per group; codes are made exact (HSUB2 128) so that the
//             rounding of A is not amplified; the bias term is one more MMA whose B columns are the three
//             per-key biases (exact in bf16).
// Fragment construction without ldmatrix: the k index of an MMA is a free permutation as long as both operands
// agree, and so is the output-column -> dim map, as long as one MMA stays inside one quantisation group:
//   * QK^T: lane (g,t) takes the 8 consecutive dims 32grp+8t..+7 of its key row for the two k-steps of a group
//     -> one 32-bit load of codes per (row, group); the Q fragments are loaded with the same permutation;
//   * PV  : column n=g of d-tile (grp, j) is dim 32grp+4g+j, so a lane needs 4 consecutive dims per (key, group)
//     -> one 16-bit load; the accumulators are un-permuted when they are written out.
// Differences from bf16(code*s+b)-then-MMA (the oracle's order) are fp32-vs-bf16 rounding of single elements,
// ~2^-9 relative per term, well inside the 2e-2 budget (tests/test_kernels_gpu.py::test_decode_attention_q4_*).
// bf16 pages (the partial last prompt page and generated tokens) go through the same fragment layout
// with plain shared-memory loads.
//
// Two orientations of the same arithmetic (template parameter HI):
//   * HI  (9..16 query rows: beam / constrained steps): queries are the MMA M dimension as described above;
//   * !HI (1..8 query rows: the decode step): TRANSPOSED — keys are the M dimension (S^T = K Q^T, O^T = V^T P^T), the
//     <= 8 query rows are N. Every MMA row is a live key instead of 8 of 16 rows being padding: half the MMAs, half the
//     accumulator registers (24 instead of 48 for O -> 4 CTAs per SM instead of 3), and the per-key scales sit on
//     accumulator ROWS. P^T reaches the B operand through movmatrix.trans (one per 8x8 block), softmax statistics are
//     per accumulator COLUMN (reduced over the 8 g-lanes).
#include "attn_common.cuh"
#include "../../include/phi3_b200.h"

#define Q4_D 96
#define QA_STAGES 6          // phase A: quantised pages, 8 KB stages
#define QA_STAGE 8192
#define QB_STAGES 2          // phase B: bf16 pages + the new tokens, 24 KB stages (same shared memory)

// two code nibbles (bits 0-3 and 16-19) -> bf16x2 (128+n_lo, 128+n_hi), exact; ONE LOP3: the mask must live in a register
// (two immediates do not fit one instruction; ptxas would emit AND + OR)
__device__ __forceinline__ uint32_t nib2bf(uint32_t x, uint32_t mask) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, 0x43004300, 0xEA;" : "=r"(d) : "r"(x), "r"(mask));
    return d;
}
__device__ __forceinline__ float meta_scale(uint32_t w) { return __uint_as_float(w << 16); }           // low half: scale (bf16)
__device__ __forceinline__ float meta_bias(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }    // high half: bias (bf16)

__device__ __forceinline__ uint32_t movm_trans(uint32_t x) {           // 8x8 b16 transpose across the warp
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(x));
    return d;
}

template <bool HI>
__global__ void __launch_bounds__(128, HI ? 3 : 4) attn_decode_q4_kernel(AttnParams p) {
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
    const uint32_t nmask = 0x000F000Fu + (uint32_t)p.zero;              // a run-time value, so that it stays in a register
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

    // Q fragments. For group grp the lane owns dims d0..d0+7, d0 = 32grp+8t; the MMA k index is permuted so that the code
    // registers are the nibble pairs one shift apart: k-step 2grp pairs dims (d0,d0+4 | d0+1,d0+5), k-step 2grp+1 pairs
    // (d0+2,d0+6 | d0+3,d0+7). HI: A fragments of rows g / g+8; !HI: B fragments of query row g.
    constexpr int NO = HI ? D / 8 : D / 16;                             // O accumulator tiles
    uint32_t qa[HI ? D / 16 : 1][4];                                    // HI
    uint32_t qb[HI ? 1 : D / 16][2];                                    // !HI
    float qs[3][2];                                                     // sum of q over each 32-dim group: HI rows g / g+8; !HI query rows 2t / 2t+1
    float o[NO][4], ob[4] = {0.f, 0.f, 0.f, 0.f};                       // ob: sum_key (P*bias - 128*P*scale)[key][grp], one column (HI) / row (!HI) per group
#pragma unroll
    for (int dt = 0; dt < NO; dt++)
#pragma unroll
        for (int j = 0; j < 4; j++) o[dt][j] = 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};             // HI: rows g / g+8; !HI: query rows 2t / 2t+1
    // operand that sums the scaled-P fragment of group grp into column / row grp of ob, times -128 (exact in bf16)
    const uint32_t kneg[3] = {g == 0 ? 0xC300C300u : 0u, g == 1 ? 0xC300C300u : 0u, g == 2 ? 0xC300C300u : 0u};

    bool q_loaded = false;
    auto load_q = [&]() {
#pragma unroll
        for (int grp = 0; grp < 3; grp++) {
            const uint32_t* lo = reinterpret_cast<const uint32_t*>(sQ + g * (D * 2) + (32 * grp + 8 * t) * 2);
            const uint32_t l0 = lo[0], l1 = lo[1], l2 = lo[2], l3 = lo[3];
            const float2 a0 = unpack_bf16(l0), a1 = unpack_bf16(l1), a2 = unpack_bf16(l2), a3 = unpack_bf16(l3);
            float a = ((a0.x + a0.y) + (a1.x + a1.y)) + ((a2.x + a2.y) + (a3.x + a3.y));
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);                    // sum over group grp of query row g
            if constexpr (HI) {
                const uint32_t* hi = reinterpret_cast<const uint32_t*>(sQ + (g + 8) * (D * 2) + (32 * grp + 8 * t) * 2);
                const uint32_t h0 = hi[0], h1 = hi[1], h2 = hi[2], h3 = hi[3];
                qa[2 * grp][0] = __byte_perm(l0, l2, 0x5410); qa[2 * grp][2] = __byte_perm(l0, l2, 0x7632);
                qa[2 * grp + 1][0] = __byte_perm(l1, l3, 0x5410); qa[2 * grp + 1][2] = __byte_perm(l1, l3, 0x7632);
                qa[2 * grp][1] = __byte_perm(h0, h2, 0x5410); qa[2 * grp][3] = __byte_perm(h0, h2, 0x7632);
                qa[2 * grp + 1][1] = __byte_perm(h1, h3, 0x5410); qa[2 * grp + 1][3] = __byte_perm(h1, h3, 0x7632);
                const float2 c0 = unpack_bf16(h0), c1 = unpack_bf16(h1), c2 = unpack_bf16(h2), c3 = unpack_bf16(h3);
                float c = ((c0.x + c0.y) + (c1.x + c1.y)) + ((c2.x + c2.y) + (c3.x + c3.y));
                c += __shfl_xor_sync(0xffffffffu, c, 1);
                c += __shfl_xor_sync(0xffffffffu, c, 2);
                qs[grp][0] = a; qs[grp][1] = c;
            } else {
                qb[2 * grp][0] = __byte_perm(l0, l2, 0x5410); qb[2 * grp][1] = __byte_perm(l0, l2, 0x7632);
                qb[2 * grp + 1][0] = __byte_perm(l1, l3, 0x5410); qb[2 * grp + 1][1] = __byte_perm(l1, l3, 0x7632);
                qs[grp][0] = __shfl_sync(0xffffffffu, a, 8 * t);        // query row 2t lives in lanes g = 2t
                qs[grp][1] = __shfl_sync(0xffffffffu, a, 8 * t + 4);
            }
        }
        q_loaded = true;
    };

    // ================= HI: queries are MMA rows ==========================================================================
    auto process_hi = [&](const uint8_t* st, int it, bool quant, bool present) {
        if (present && warp != 0) return;
        const int key0 = present ? 0 : warp * 16;                       // this warp's 16 keys inside the tile
        // ---- S = Q K^T
        float s[2][4];
        if (quant) {
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                const uint8_t* row = st + (key0 + nt * 8 + g) * (D / 2);
                float acc[3][4];
#pragma unroll
                for (int grp = 0; grp < 3; grp++) {
                    acc[grp][0] = acc[grp][1] = acc[grp][2] = acc[grp][3] = 0.f;
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(row + grp * 16 + 4 * t);   // nibbles n0..n7 = dims d0..d0+7
                    mma_bf16_16816(acc[grp], qa[2 * grp], nib2bf(w, nmask), nib2bf(w >> 4, nmask));          // (n0,n4) (n1,n5)
                    mma_bf16_16816(acc[grp], qa[2 * grp + 1], nib2bf(w >> 8, nmask), nib2bf(w >> 12, nmask)); // (n2,n6) (n3,n7)
                }
                // accumulator columns 2t, 2t+1 = keys key0+nt*8+2t(+1): their 3 (scale, bias) words are 24 contiguous bytes
                const uint2* mp = reinterpret_cast<const uint2*>(st + 2 * QC) + (key0 + nt * 8 + 2 * t) * 3 / 2;
                const uint2 m0 = mp[0], m1 = mp[1], m2 = mp[2];
                const uint32_t mw[2][3] = {{m0.x, m0.y, m1.x}, {m1.y, m2.x, m2.y}};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float v = 0.f;
#pragma unroll
                    for (int grp = 0; grp < 3; grp++) {
                        const float sc = meta_scale(mw[e & 1][grp]);
                        v = fmaf(sc, acc[grp][e], v);
                        v = fmaf(fmaf(sc, -128.f, meta_bias(mw[e & 1][grp])), qs[grp][e >> 1], v);
                    }
                    s[nt][e] = v;
                }
            }
        } else {
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
                const uint8_t* row = st + (key0 + nt * 8 + g) * (D * 2);
#pragma unroll
                for (int grp = 0; grp < 3; grp++) {
                    const uint4 kv = *reinterpret_cast<const uint4*>(row + (32 * grp + 8 * t) * 2);
                    mma_bf16_16816(s[nt], qa[2 * grp], __byte_perm(kv.x, kv.z, 0x5410), __byte_perm(kv.x, kv.z, 0x7632));
                    mma_bf16_16816(s[nt], qa[2 * grp + 1], __byte_perm(kv.y, kv.w, 0x5410), __byte_perm(kv.y, kv.w, 0x7632));
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
            for (int dt = 0; dt < NO; dt++) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
            ob[0] *= c0; ob[1] *= c0; ob[2] *= c1; ob[3] *= c1;
        }
        // ---- O += P V ; column g of d-tile (grp, j) is dim 32grp+4g+j; MMA k rows = keys 2t, 2t+1, 2t+8, 2t+9
        const int R0 = key0 + 2 * t;
        if (quant) {
            const uint8_t* vc = st + QC + R0 * (D / 2) + 2 * g;
            const uint2* mp = reinterpret_cast<const uint2*>(st + 2 * QC + QM) + R0 * 3 / 2;
            const uint2 m0 = mp[0], m1 = mp[1], m2 = mp[2], m3 = mp[12], m4 = mp[13], m5 = mp[14];   // +8 keys = +12 uint2
            const uint32_t mw[4][3] = {{m0.x, m0.y, m1.x}, {m1.y, m2.x, m2.y}, {m3.x, m3.y, m4.x}, {m4.y, m5.x, m5.y}};
            {   // bias term: B column n = group n (n < 3), rows = the 4 keys' biases (exact bf16)
                const uint32_t a[4] = {pack_bf16(s[0][0], s[0][1]), pack_bf16(s[0][2], s[0][3]),
                                       pack_bf16(s[1][0], s[1][1]), pack_bf16(s[1][2], s[1][3])};
                uint32_t wsel[4];
#pragma unroll
                for (int k = 0; k < 4; k++) wsel[k] = (g == 0) ? mw[k][0] : (g == 1) ? mw[k][1] : mw[k][2];
                const uint32_t b0 = (g < 3) ? __byte_perm(wsel[0], wsel[1], 0x7632) : 0u;
                const uint32_t b1 = (g < 3) ? __byte_perm(wsel[2], wsel[3], 0x7632) : 0u;
                mma_bf16_16816(ob, a, b0, b1);
            }
#pragma unroll
            for (int grp = 0; grp < 3; grp++) {
                const float s0 = meta_scale(mw[0][grp]), s1 = meta_scale(mw[1][grp]), s2 = meta_scale(mw[2][grp]), s3 = meta_scale(mw[3][grp]);
                const uint32_t a[4] = {pack_bf16(s[0][0] * s0, s[0][1] * s1), pack_bf16(s[0][2] * s0, s[0][3] * s1),
                                       pack_bf16(s[1][0] * s2, s[1][1] * s3), pack_bf16(s[1][2] * s2, s[1][3] * s3)};
                const uint32_t h0 = *reinterpret_cast<const uint16_t*>(vc + grp * 16);
                const uint32_t h1 = *reinterpret_cast<const uint16_t*>(vc + (D / 2) + grp * 16);
                const uint32_t h2 = *reinterpret_cast<const uint16_t*>(vc + 8 * (D / 2) + grp * 16);
                const uint32_t h3 = *reinterpret_cast<const uint16_t*>(vc + 9 * (D / 2) + grp * 16);
                const uint32_t w01 = __byte_perm(h0, h1, 0x5410), w23 = __byte_perm(h2, h3, 0x5410);
                mma_bf16_16816(ob, a, kneg[grp], kneg[grp]);            // -128 * sum_key A_grp: undoes the +128 of the codes exactly
#pragma unroll
                for (int j = 0; j < 4; j++)
                    mma_bf16_16816(o[grp * 4 + j], a, nib2bf(w01 >> (4 * j), nmask), nib2bf(w23 >> (4 * j), nmask));
            }
        } else {
            const uint32_t a[4] = {pack_bf16(s[0][0], s[0][1]), pack_bf16(s[0][2], s[0][3]),
                                   pack_bf16(s[1][0], s[1][1]), pack_bf16(s[1][2], s[1][3])};
            const uint8_t* vt = st + 64 * D * 2 + R0 * (D * 2) + 8 * g;
#pragma unroll
            for (int grp = 0; grp < 3; grp++) {
                const uint2 x0 = *reinterpret_cast<const uint2*>(vt + grp * 64);
                const uint2 x1 = *reinterpret_cast<const uint2*>(vt + (D * 2) + grp * 64);
                const uint2 x2 = *reinterpret_cast<const uint2*>(vt + 8 * (D * 2) + grp * 64);
                const uint2 x3 = *reinterpret_cast<const uint2*>(vt + 9 * (D * 2) + grp * 64);
                mma_bf16_16816(o[grp * 4 + 0], a, __byte_perm(x0.x, x1.x, 0x5410), __byte_perm(x2.x, x3.x, 0x5410));
                mma_bf16_16816(o[grp * 4 + 1], a, __byte_perm(x0.x, x1.x, 0x7632), __byte_perm(x2.x, x3.x, 0x7632));
                mma_bf16_16816(o[grp * 4 + 2], a, __byte_perm(x0.y, x1.y, 0x5410), __byte_perm(x2.y, x3.y, 0x5410));
                mma_bf16_16816(o[grp * 4 + 3], a, __byte_perm(x0.y, x1.y, 0x7632), __byte_perm(x2.y, x3.y, 0x7632));
            }
        }
    };

    // ================= !HI: keys are MMA rows (S^T = K Q^T, O^T = V^T P^T), query rows 0..7 are the columns =============
    // accumulator element e of a lane: key row g (e < 2) / g+8 (e >= 2), query row 2t + (e & 1)
    auto process_lo = [&](const uint8_t* st, int it, bool quant, bool present) {
        if (present && warp != 0) return;
        const int key0 = present ? 0 : warp * 16;
        float s[4];
        if (quant) {
            const uint8_t* ka = st + (key0 + g) * (D / 2) + 4 * t;      // key g; key g+8 is 8 rows further
            float acc[3][4];
#pragma unroll
            for (int grp = 0; grp < 3; grp++) {
                acc[grp][0] = acc[grp][1] = acc[grp][2] = acc[grp][3] = 0.f;
                const uint32_t wa = *reinterpret_cast<const uint32_t*>(ka + grp * 16);
                const uint32_t wb = *reinterpret_cast<const uint32_t*>(ka + 8 * (D / 2) + grp * 16);
                const uint32_t a1[4] = {nib2bf(wa, nmask), nib2bf(wb, nmask), nib2bf(wa >> 4, nmask), nib2bf(wb >> 4, nmask)};
                const uint32_t a2[4] = {nib2bf(wa >> 8, nmask), nib2bf(wb >> 8, nmask), nib2bf(wa >> 12, nmask), nib2bf(wb >> 12, nmask)};
                mma_bf16_16816(acc[grp], a1, qb[2 * grp][0], qb[2 * grp][1]);
                mma_bf16_16816(acc[grp], a2, qb[2 * grp + 1][0], qb[2 * grp + 1][1]);
            }
            const uint32_t* ma = reinterpret_cast<const uint32_t*>(st + 2 * QC) + (key0 + g) * 3;
            s[0] = s[1] = s[2] = s[3] = 0.f;
#pragma unroll
            for (int grp = 0; grp < 3; grp++) {
                const uint32_t wa = ma[grp], wb = ma[24 + grp];          // keys g, g+8
                const float sa = meta_scale(wa), sb = meta_scale(wb);
                const float ca = fmaf(sa, -128.f, meta_bias(wa)), cb = fmaf(sb, -128.f, meta_bias(wb));
                s[0] = fmaf(sa, acc[grp][0], fmaf(ca, qs[grp][0], s[0]));
                s[1] = fmaf(sa, acc[grp][1], fmaf(ca, qs[grp][1], s[1]));
                s[2] = fmaf(sb, acc[grp][2], fmaf(cb, qs[grp][0], s[2]));
                s[3] = fmaf(sb, acc[grp][3], fmaf(cb, qs[grp][1], s[3]));
            }
        } else {
            s[0] = s[1] = s[2] = s[3] = 0.f;
            const uint8_t* ka = st + (key0 + g) * (D * 2) + 16 * t;
#pragma unroll
            for (int grp = 0; grp < 3; grp++) {
                const uint4 va = *reinterpret_cast<const uint4*>(ka + grp * 64);
                const uint4 vb = *reinterpret_cast<const uint4*>(ka + 8 * (D * 2) + grp * 64);
                const uint32_t a1[4] = {__byte_perm(va.x, va.z, 0x5410), __byte_perm(vb.x, vb.z, 0x5410),
                                        __byte_perm(va.x, va.z, 0x7632), __byte_perm(vb.x, vb.z, 0x7632)};
                const uint32_t a2[4] = {__byte_perm(va.y, va.w, 0x5410), __byte_perm(vb.y, vb.w, 0x5410),
                                        __byte_perm(va.y, va.w, 0x7632), __byte_perm(vb.y, vb.w, 0x7632)};
                mma_bf16_16816(s, a1, qb[2 * grp][0], qb[2 * grp][1]);
                mma_bf16_16816(s, a2, qb[2 * grp + 1][0], qb[2 * grp + 1][1]);
            }
        }
        // ---- mask (boundary tiles only)
        const int j0 = present ? past : (n_lo + it) * 64 + warp * 16;
        if (present || j0 < kv0 || j0 + 16 > past) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int j = j0 + g + ((e & 2) ? 8 : 0);
                bool ok;
                if (present) { const int qi = past + 2 * t + (e & 1); ok = (j >= kv0) && (j < s_total) && (j <= qi); }
                else ok = (j >= kv0) && (j < past);
                if (!ok) s[e] = -INFINITY;
            }
        }
        // ---- online softmax: statistics per accumulator column (query row), reduced over the 8 g-lanes
        float mx0 = fmaxf(s[0], s[2]), mx1 = fmaxf(s[1], s[3]);
#pragma unroll
        for (int sh = 4; sh < 32; sh <<= 1) {
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, sh));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, sh));
        }
        const float mn0 = fmaxf(m[0], mx0), mn1 = fmaxf(m[1], mx1);
        const float mu0 = (mn0 == -INFINITY) ? 0.f : mn0 * p.scale_log2, mu1 = (mn1 == -INFINITY) ? 0.f : mn1 * p.scale_log2;
        const float c0 = ex2_approx(m[0] * p.scale_log2 - mu0), c1 = ex2_approx(m[1] * p.scale_log2 - mu1);
        m[0] = mn0; m[1] = mn1;
        s[0] = ex2_approx(fmaf(s[0], p.scale_log2, -mu0)); s[1] = ex2_approx(fmaf(s[1], p.scale_log2, -mu1));
        s[2] = ex2_approx(fmaf(s[2], p.scale_log2, -mu0)); s[3] = ex2_approx(fmaf(s[3], p.scale_log2, -mu1));
        l[0] = fmaf(l[0], c0, s[0] + s[2]);                             // this lane's two keys; summed over g at the end
        l[1] = fmaf(l[1], c1, s[1] + s[3]);
        if (c0 != 1.f || c1 != 1.f) {
#pragma unroll
            for (int dt = 0; dt < NO; dt++) { o[dt][0] *= c0; o[dt][1] *= c1; o[dt][2] *= c0; o[dt][3] *= c1; }
            ob[0] *= c0; ob[1] *= c1; ob[2] *= c0; ob[3] *= c1;
        }
        // ---- O^T += V^T P^T. A rows: tile (grp, i) row g = dim 32grp+4g+2i, row g+8 = that dim + 1; k = keys 2t,2t+1 | 2t+8,2t+9.
        //      B = P^T: the accumulator halves (keys g / g+8 x query rows 2t,2t+1) transposed 8x8 by movmatrix
        const int R0 = key0 + 2 * t;
        if (quant) {
            const uint32_t* vm = reinterpret_cast<const uint32_t*>(st + 2 * QC + QM);
            {   // bias term: A row r = group r (r < 3) holds the 4 keys' biases of that group (exact bf16)
                const uint32_t b0 = movm_trans(pack_bf16(s[0], s[1])), b1 = movm_trans(pack_bf16(s[2], s[3]));
                const uint32_t* mr = vm + R0 * 3 + min(g, 2);
                const uint32_t a[4] = {(g < 3) ? __byte_perm(mr[0], mr[3], 0x7632) : 0u, 0u,
                                       (g < 3) ? __byte_perm(mr[24], mr[27], 0x7632) : 0u, 0u};
                mma_bf16_16816(ob, a, b0, b1);
            }
            const uint32_t* mk = vm + (key0 + g) * 3;                   // scales of this lane's accumulator rows: keys g, g+8
            const uint8_t* vc = st + QC + R0 * (D / 2) + 2 * g;
#pragma unroll
            for (int grp = 0; grp < 3; grp++) {
                const float sa = meta_scale(mk[grp]), sb = meta_scale(mk[24 + grp]);
                const uint32_t b0 = movm_trans(pack_bf16(s[0] * sa, s[1] * sa)), b1 = movm_trans(pack_bf16(s[2] * sb, s[3] * sb));
                const uint32_t h0 = *reinterpret_cast<const uint16_t*>(vc + grp * 16);
                const uint32_t h1 = *reinterpret_cast<const uint16_t*>(vc + (D / 2) + grp * 16);
                const uint32_t h2 = *reinterpret_cast<const uint16_t*>(vc + 8 * (D / 2) + grp * 16);
                const uint32_t h3 = *reinterpret_cast<const uint16_t*>(vc + 9 * (D / 2) + grp * 16);
                const uint32_t w01 = __byte_perm(h0, h1, 0x5410), w23 = __byte_perm(h2, h3, 0x5410);
                const uint32_t an[4] = {kneg[grp], 0u, kneg[grp], 0u};  // -128 * sum_key (P*scale): undoes the +128 of the codes exactly
                mma_bf16_16816(ob, an, b0, b1);
                const uint32_t a1[4] = {nib2bf(w01, nmask), nib2bf(w01 >> 4, nmask), nib2bf(w23, nmask), nib2bf(w23 >> 4, nmask)};
                const uint32_t a2[4] = {nib2bf(w01 >> 8, nmask), nib2bf(w01 >> 12, nmask), nib2bf(w23 >> 8, nmask), nib2bf(w23 >> 12, nmask)};
                mma_bf16_16816(o[grp * 2], a1, b0, b1);
                mma_bf16_16816(o[grp * 2 + 1], a2, b0, b1);
            }
        } else {
            const uint32_t b0 = movm_trans(pack_bf16(s[0], s[1])), b1 = movm_trans(pack_bf16(s[2], s[3]));
            const uint8_t* vt = st + 64 * D * 2 + R0 * (D * 2) + 8 * g;
#pragma unroll
            for (int grp = 0; grp < 3; grp++) {
                const uint2 x0 = *reinterpret_cast<const uint2*>(vt + grp * 64);
                const uint2 x1 = *reinterpret_cast<const uint2*>(vt + (D * 2) + grp * 64);
                const uint2 x2 = *reinterpret_cast<const uint2*>(vt + 8 * (D * 2) + grp * 64);
                const uint2 x3 = *reinterpret_cast<const uint2*>(vt + 9 * (D * 2) + grp * 64);
                const uint32_t a1[4] = {__byte_perm(x0.x, x1.x, 0x5410), __byte_perm(x0.x, x1.x, 0x7632),
                                        __byte_perm(x2.x, x3.x, 0x5410), __byte_perm(x2.x, x3.x, 0x7632)};
                const uint32_t a2[4] = {__byte_perm(x0.y, x1.y, 0x5410), __byte_perm(x0.y, x1.y, 0x7632),
                                        __byte_perm(x2.y, x3.y, 0x5410), __byte_perm(x2.y, x3.y, 0x7632)};
                mma_bf16_16816(o[grp * 2], a1, b0, b1);
                mma_bf16_16816(o[grp * 2 + 1], a2, b0, b1);
            }
        }
    };
    auto process = [&](const uint8_t* st, int it, bool quant, bool present) {
        if constexpr (HI) process_hi(st, it, quant, present); else process_lo(st, it, quant, present);
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
    if constexpr (HI) {
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
        for (int grp = 0; grp < 3; grp++) {
            // the bias accumulator of (row, grp) lives in lane (g, grp>>1), element grp&1 (+2 for row g+8)
            const float bl = __shfl_sync(0xffffffffu, (grp & 1) ? ob[1] : ob[0], 4 * g + (grp >> 1));
            const float bh = __shfl_sync(0xffffffffu, (grp & 1) ? ob[3] : ob[2], 4 * g + (grp >> 1));
#pragma unroll
            for (int j = 0; j < 4; j++) {                               // accumulator column 2t(+1) of tile (grp, j) = dim 32grp+8t(+4)+j
                const int dt = grp * 4 + j, d0 = 32 * grp + 8 * t + j;
                sm_o[(warp * 16 + g) * D + d0] = o[dt][0] + bl;
                sm_o[(warp * 16 + g) * D + d0 + 4] = o[dt][1] + bl;
                sm_o[(warp * 16 + g + 8) * D + d0] = o[dt][2] + bh;
                sm_o[(warp * 16 + g + 8) * D + d0 + 4] = o[dt][3] + bh;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int sh = 4; sh < 32; sh <<= 1) l[i] += __shfl_xor_sync(0xffffffffu, l[i], sh);
        if (g == 0) {
            sm_m[warp * 16 + 2 * t] = m[0] * p.scale_log2; sm_m[warp * 16 + 2 * t + 1] = m[1] * p.scale_log2;
            sm_l[warp * 16 + 2 * t] = l[0]; sm_l[warp * 16 + 2 * t + 1] = l[1];
        }
#pragma unroll
        for (int grp = 0; grp < 3; grp++) {
            // the bias accumulator row grp lives in lanes g = grp: elements 0, 1 = query rows 2t, 2t+1
            const float b0 = __shfl_sync(0xffffffffu, ob[0], 4 * grp + t), b1 = __shfl_sync(0xffffffffu, ob[1], 4 * grp + t);
#pragma unroll
            for (int i = 0; i < 2; i++) {                               // tile (grp, i): rows g / g+8 = dims 32grp+4g+2i (+1)
                const int dt = grp * 2 + i, d0 = 32 * grp + 4 * g + 2 * i;
                sm_o[(warp * 16 + 2 * t) * D + d0] = o[dt][0] + b0;
                sm_o[(warp * 16 + 2 * t + 1) * D + d0] = o[dt][1] + b1;
                sm_o[(warp * 16 + 2 * t) * D + d0 + 1] = o[dt][2] + b0;
                sm_o[(warp * 16 + 2 * t + 1) * D + d0 + 1] = o[dt][3] + b1;
            }
        }
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
        cudaError_t e = cudaFuncSetAttribute(attn_decode_q4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_decode_q4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        P3_CHECK_ARG(e == cudaSuccess, "attention_decode_q4: smem attribute: %s", cudaGetErrorString(e));
        set = true;
    }
    if (p.L > 8) p3_launch_pdl(attn_decode_q4_kernel<true>, grid, dim3(128), (size_t)smem, st, p);
    else p3_launch_pdl(attn_decode_q4_kernel<false>, grid, dim3(128), (size_t)smem, st, p);
    P3_CHECK_LAUNCH("attention_decode_q4");
    return 0;
}
