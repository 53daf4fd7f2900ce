// Attention kernels.
//  * attn_prefill_kernel : flash-style tiled softmax(QK^T)V for many query rows (prefill, ViT).
//    Replaces the explicit scores/softmax/PV of phi.py:454-457 and the dense Mask4D of
//    phi.py:550-563 (predicate: causal + per-row left-pad start) and
//    mx.fast.scaled_dot_product_attention of phi.py:148 (causal = 0).
//  * attn_decode_kernel  : (bf16; the 4-bit variant lives in attention_q4.cu) memory-bound split-KV attention for <=16 new tokens per row over the
//    paged bf16 KV pool (phi.py:454-457 + KVCache reads phi.py:523-527,548). 16-byte coalesced
//    cp.async loads of whole 12 KB page slices into a 4-stage shared-memory ring, warp-shuffle
//    softmax reductions, tensor-core (mma.sync) dot products so L<=16 queries cost one pass.
//    (split partials are merged in-kernel by the last CTA of each (row, head) to arrive)
// Keys [0,past) are read from the paged pool, keys [past,past+L) from the freshly roped
// qkv buffer ("dual source"), which is what makes the read-only beam step of phi.py:523-527 a
// plain call with row_div = n_beam.
#include "attn_common.cuh"
#include "../../include/phi3_b200.h"

// Load a 64-key K tile and V tile (absolute key positions [64n, 64n+64)) into shared memory.
template <int D, int THREADS>
__device__ __forceinline__ void load_kv_tile(const AttnParams& p, int b, int kvh, int n, int s_total, uint32_t sk, uint32_t sv) {
    constexpr int CPR = D / 8;
    const int crow = b / p.row_div;
    const bf16 *kpage = nullptr, *vpage = nullptr;
    if (n * P3_PAGE < p.past) {
        int page = p.block_table[(size_t)crow * p.bt_stride + n];
        kpage = kv_tile_ptr(p.pool, page, 0, kvh, p.n_kv, D);
        vpage = kv_tile_ptr(p.pool, page, 1, kvh, p.n_kv, D);
    }
    for (int idx = threadIdx.x; idx < 64 * CPR; idx += THREADS) {
        int r = idx / CPR, c = idx % CPR;
        int j = n * P3_PAGE + r;
        const bf16 *ks, *vs;
        int bytes = 16;
        if (j < p.past) {
            ks = kpage + idx * 8; vs = vpage + idx * 8;
        } else if (j < s_total) {
            size_t tok = (size_t)b * p.L + (j - p.past);
            ks = p.k + tok * p.ldk + kvh * D + c * 8;
            vs = p.v + tok * p.ldv + kvh * D + c * 8;
        } else {
            ks = p.k; vs = p.v; bytes = 0;                      // zero fill
        }
        cp_async16(sk + tile_off<D>(r, c), ks, bytes);
        cp_async16(sv + tile_off<D>(r, c), vs, bytes);
    }
}

// One warp: S(16 x NK*8 keys) = Q K^T for keys [key0, key0 + 8*NK) of the tile at sk.
template <int D, int NK>
__device__ __forceinline__ void qk_mma(const uint32_t (*qf)[4], uint32_t sk, int key0, float (*s)[4], int lane) {
#pragma unroll
    for (int nt = 0; nt < NK; nt++)
#pragma unroll
        for (int j = 0; j < 4; j++) s[nt][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < D / 16; ks++) {
#pragma unroll
        for (int np = 0; np < NK / 2; np++) {
            int r = key0 + np * 16 + (lane >> 4) * 8 + (lane & 7);
            int c = ks * 2 + ((lane >> 3) & 1);
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(b0, b1, b2, b3, sk + tile_off<D>(r, c));
            mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
            mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
        }
    }
}

// One warp: O(16 x D) += P(16 x NK*8 keys) V. P is rounded to bf16 for the tensor cores (the
// reference runs softmax.V in fp32, phi.py:454-457; a bf16 hi/lo split of P was measured to change
// nothing: bf16 pipelines decorrelate at their noise floor regardless).
template <int D, int NK>
__device__ __forceinline__ void pv_mma(const float (*s)[4], uint32_t sv, int key0, float (*o)[4], int lane) {
#pragma unroll
    for (int kk = 0; kk < NK / 2; kk++) {
        uint32_t a[4] = {pack_bf16(s[2 * kk][0], s[2 * kk][1]), pack_bf16(s[2 * kk][2], s[2 * kk][3]),
                         pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
        for (int dp = 0; dp < D / 16; dp++) {
            int r = key0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            int c = dp * 2 + (lane >> 4);
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(b0, b1, b2, b3, sv + tile_off<D>(r, c));
            mma_bf16_16816(o[2 * dp], a, b0, b1);
            mma_bf16_16816(o[2 * dp + 1], a, b2, b3);
        }
    }
}

// Online-softmax update for one warp's 16 x (NK*8) score block. Rows g (idx 0) and g+8 (idx 1).
template <int D, int NK>
__device__ __forceinline__ void softmax_update(float (*s)[4], float* m, float* l, float (*o)[4]) {
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NK; nt++) {
        mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
        mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    }
    float corr[2], mu[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        float mn = fmaxf(m[i], mx[i]);
        mu[i] = (mn == -INFINITY) ? 0.f : mn;
        corr[i] = exp2f(m[i] - mu[i]);                        // m = -inf -> 0
        m[i] = mn;
        l[i] *= corr[i];
    }
#pragma unroll
    for (int nt = 0; nt < NK; nt++) {
        s[nt][0] = exp2f(s[nt][0] - mu[0]); s[nt][1] = exp2f(s[nt][1] - mu[0]);
        s[nt][2] = exp2f(s[nt][2] - mu[1]); s[nt][3] = exp2f(s[nt][3] - mu[1]);
        l[0] += s[nt][0] + s[nt][1];
        l[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int dt = 0; dt < D / 8; dt++) {
        o[dt][0] *= corr[0]; o[dt][1] *= corr[0]; o[dt][2] *= corr[1]; o[dt][3] *= corr[1];
    }
}

// ------------------------------------------------------------------------------------------
// prefill / ViT: 128 query rows per CTA (8 warps x 16 rows), 64-key tiles, 3-stage cp.async ring.
// Per-thread copy slots are hoisted out of the tile loop for the two common cases (tile entirely in
// the paged pool / entirely in the fresh qkv buffer); masking runs only on boundary tiles; a warp
// skips tiles that lie entirely above its causal diagonal; exp2 is one FFMA + one MUFU.
// ------------------------------------------------------------------------------------------
#define PF_STAGES 3
#define PF_THREADS 256

template <int D>
__global__ void __launch_bounds__(PF_THREADS, 2) attn_prefill_kernel(AttnParams p) {
    constexpr int CPR = D / 8, TILE = 64 * D * 2, NSLOT = 64 * CPR / PF_THREADS;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sq = smem_u32(smem), skv = sq + 2 * TILE;   // Q: 128 rows; skv: [PF_STAGES][K,V]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int i0 = (gridDim.x - 1 - blockIdx.x) * 128, h = blockIdx.y, b = blockIdx.z;   // heavy (late) row blocks first
    const int kvh = h / (p.n_heads / p.n_kv);
    const int s_total = p.past + p.L;
    const int crow = b / p.row_div;
    const int kv0 = p.kv_start ? p.kv_start[crow] : 0;

    for (int idx = tid; idx < 128 * CPR; idx += PF_THREADS) {
        int r = idx / CPR, c = idx % CPR;
        int i = i0 + r;
        const bf16* src = p.q + ((size_t)b * p.L + (i < p.L ? i : 0)) * p.ldq + h * D + c * 8;
        cp_async16(sq + tile_off<D>(r, c), src, i < p.L ? 16 : 0);
    }
    const int n_begin = kv0 / 64;
    int n_end = (s_total + 63) / 64;
    if (p.causal) n_end = min(n_end, (p.past + min(i0 + 127, p.L - 1)) / 64 + 1);
    const int n_tiles = max(n_end - n_begin, 0);

    // hoisted copy slots
    uint32_t soff[NSLOT];
    int goff[NSLOT];                                            // element offset inside the fresh k/v rows
#pragma unroll
    for (int i = 0; i < NSLOT; i++) {
        int idx = tid + PF_THREADS * i, r = idx / CPR, c = idx % CPR;
        soff[i] = tile_off<D>(r, c);
        goff[i] = r * (int)p.ldk + c * 8;
    }
    const size_t head_elems = (size_t)P3_PAGE * D, page_elems = 2 * (size_t)p.n_kv * head_elems;
    const int32_t* bt = p.block_table ? p.block_table + (size_t)crow * p.bt_stride : nullptr;

    auto issue = [&](int n, int stage) {
        const uint32_t sk = skv + stage * 2 * TILE, sv = sk + TILE;
        const int j0 = n * 64;
        if (j0 + 64 <= p.past) {                                // whole tile from the paged pool
            const bf16* kp = p.pool + (size_t)bt[n] * page_elems + (size_t)kvh * head_elems + tid * 8;
            const bf16* vp = kp + (size_t)p.n_kv * head_elems;
#pragma unroll
            for (int i = 0; i < NSLOT; i++) {
                cp_async16(sk + soff[i], kp + i * (PF_THREADS * 8));
                cp_async16(sv + soff[i], vp + i * (PF_THREADS * 8));
            }
        } else if (j0 >= p.past && j0 + 64 <= s_total && p.ldk == p.ldv) {   // whole tile from the fresh qkv rows
            const size_t tok0 = (size_t)b * p.L + (j0 - p.past);
            const bf16* kp = p.k + tok0 * p.ldk + kvh * D;
            const bf16* vp = p.v + tok0 * p.ldv + kvh * D;
#pragma unroll
            for (int i = 0; i < NSLOT; i++) {
                cp_async16(sk + soff[i], kp + goff[i]);
                cp_async16(sv + soff[i], vp + goff[i]);
            }
        } else {
            load_kv_tile<D, PF_THREADS>(p, b, kvh, n, s_total, sk, sv);
        }
    };
#pragma unroll
    for (int i = 0; i < PF_STAGES - 1; i++) {
        if (i < n_tiles) issue(n_begin + i, i);
        cp_async_commit();                                      // group 0 also carries the Q tile
    }

    uint32_t qf[D / 16][4];
    float o[D / 8][4];
#pragma unroll
    for (int dt = 0; dt < D / 8; dt++)
#pragma unroll
        for (int j = 0; j < 4; j++) o[dt][j] = 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    const int qi0 = p.past + i0 + warp * 16 + g, qi1 = qi0 + 8;   // absolute query positions of rows g, g+8
    const int q_hi = p.past + i0 + warp * 16 + 15;               // last query row of this warp

    for (int itn = 0; itn < n_tiles; itn++) {
        const int n = n_begin + itn;
        cp_async_wait<PF_STAGES - 2>();
        __syncthreads();
        if (itn + PF_STAGES - 1 < n_tiles) issue(n + PF_STAGES - 1, (itn + PF_STAGES - 1) % PF_STAGES);
        cp_async_commit();
        if (itn == 0) {
#pragma unroll
            for (int ks = 0; ks < D / 16; ks++) {
                int r = warp * 16 + (lane & 15), c = ks * 2 + (lane >> 4);
                ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], sq + tile_off<D>(r, c));
            }
        }
        const int j0 = n * 64;
        if (p.causal && j0 > q_hi) continue;                    // tile entirely above this warp's diagonal
        const uint32_t sk = skv + (itn % PF_STAGES) * 2 * TILE, sv = sk + TILE;
        float s[8][4];
        qk_mma<D, 8>(qf, sk, 0, s, lane);
        const bool need_mask = (j0 < kv0) || (j0 + 63 >= s_total) || (p.causal && j0 + 63 > p.past + i0 + warp * 16);
        if (need_mask) {
#pragma unroll
            for (int nt = 0; nt < 8; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    int j = j0 + nt * 8 + 2 * t + (e & 1);
                    int qi = (e & 2) ? qi1 : qi0;
                    bool ok = (j >= kv0) && (j < s_total) && (!p.causal || j <= qi);
                    if (!ok) s[nt][e] = -INFINITY;
                }
        }
        // online softmax in the exp2 domain on raw scores (scale folded into the FFMA)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m[0], mx0), mn1 = fmaxf(m[1], mx1);
        const float mu0 = (mn0 == -INFINITY) ? 0.f : mn0 * p.scale_log2, mu1 = (mn1 == -INFINITY) ? 0.f : mn1 * p.scale_log2;
        const float c0 = ex2_approx(m[0] * p.scale_log2 - mu0), c1 = ex2_approx(m[1] * p.scale_log2 - mu1);
        m[0] = mn0; m[1] = mn1;
        float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            s[nt][0] = ex2_approx(fmaf(s[nt][0], p.scale_log2, -mu0)); s[nt][1] = ex2_approx(fmaf(s[nt][1], p.scale_log2, -mu0));
            s[nt][2] = ex2_approx(fmaf(s[nt][2], p.scale_log2, -mu1)); s[nt][3] = ex2_approx(fmaf(s[nt][3], p.scale_log2, -mu1));
            ls0 += s[nt][0] + s[nt][1]; ls1 += s[nt][2] + s[nt][3];
        }
        l[0] = fmaf(l[0], c0, ls0); l[1] = fmaf(l[1], c1, ls1);
#pragma unroll
        for (int dt = 0; dt < D / 8; dt++) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
        pv_mma<D, 8>(s, sv, 0, o, lane);
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 2; i++) {
        l[i] += __shfl_xor_sync(0xffffffffu, l[i], 1);
        l[i] += __shfl_xor_sync(0xffffffffu, l[i], 2);
    }
    const float inv0 = l[0] > 0.f ? 1.f / l[0] : 0.f, inv1 = l[1] > 0.f ? 1.f / l[1] : 0.f;
    const int r0 = i0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int dt = 0; dt < D / 8; dt++) {
        int d = dt * 8 + 2 * t;
        if (r0 < p.L)
            *reinterpret_cast<uint32_t*>(p.out + ((size_t)b * p.L + r0) * p.ldo + h * D + d) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
        if (r1 < p.L)
            *reinterpret_cast<uint32_t*>(p.out + ((size_t)b * p.L + r1) * p.ldo + h * D + d) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
    }
}

// ------------------------------------------------------------------------------------------
// decode (L <= 16 new tokens per row), split-KV over cached pages.
// grid (n_splits, n_heads, B), 128 threads. Each CTA streams its share of the row's 64-token
// pages (12 KB K + 12 KB V per head) through a 4-stage cp.async ring; warp w owns keys
// [16w,16w+16) of every tile, so a tile costs each warp 12 QK + 12 PV mma.sync. All copy and
// ldmatrix offsets are hoisted out of the tile loop, the page id of tile t+4 is fetched while
// tile t is consumed, masking runs only on boundary tiles and exp2 is a single MUFU.
// ------------------------------------------------------------------------------------------
#define DEC_STAGES 4
// timeline stamps (tools/chain_trace.py): compiled in only with -DP3_TRACE_ATTN — any extra uniform-register use in this kernel
// makes ptxas 12.9 pick the [R + UR + imm] LDGSTS form with an odd policy descriptor register (see cp_async16_stream)
#ifdef P3_TRACE_ATTN
#define ATTN_TRACE_STAMP(tr, cta, i) trace_stamp(tr, cta, i)
#define ATTN_TRACE_META(tr, cta, k) trace_meta(tr, cta, k)
#else
#define ATTN_TRACE_STAMP(tr, cta, i)
#define ATTN_TRACE_META(tr, cta, k)
#endif

template <int D>
__global__ void __launch_bounds__(128, 2) attn_decode_kernel(AttnParams p) {
    constexpr int CPR = D / 8, TILE = 64 * D * 2, NSLOT = 64 * CPR / 128;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sq = smem_u32(smem);                         // 16 x D query tile
    const uint32_t skv = sq + 16 * D * 2;                       // [DEC_STAGES][K,V]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (p.n_heads / p.n_kv);
    ATTN_TRACE_STAMP(p.trace, (int)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x), 0);
    ATTN_TRACE_META(p.trace, (int)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x), 900);
    pdl_trigger();
    // hoisted per-thread copy slots: chunk (tid + 128 i) of the contiguous 64 x D page slice
    uint32_t soff[NSLOT];
#pragma unroll
    for (int i = 0; i < NSLOT; i++) {
        int idx = tid + 128 * i;
        soff[i] = tile_off<D>(idx / CPR, idx % CPR);
    }
    const size_t head_elems = (size_t)P3_PAGE * D, page_elems = 2 * (size_t)p.n_kv * head_elems;
    const bf16* pool_k = p.pool + (size_t)kvh * head_elems + tid * 8;          // + page * page_elems
    const size_t v_delta = (size_t)p.n_kv * head_elems;

    const uint64_t pol = l2_evict_first_policy();
    auto issue_cached = [&](int it, int page) {
        const uint32_t sk = skv + (it % DEC_STAGES) * 2 * TILE;
        const bf16* kp = pool_k + (size_t)page * page_elems;
#pragma unroll
        for (int i = 0; i < NSLOT; i++) {
            if (D == 96) {          // KV pages are read once per step: L2 evict-first
                cp_async16_stream(sk + soff[i], kp + i * 1024, pol);
                cp_async16_stream(sk + TILE + soff[i], kp + v_delta + i * 1024, pol);
            } else {                // (ptxas 12.9 emits an unencodable LDGSTS descriptor pair for the D=64 instance)
                cp_async16(sk + soff[i], kp + i * 1024);
                cp_async16(sk + TILE + soff[i], kp + v_delta + i * 1024);
            }
        }
    };
    // Before the dependency wait: fill the ring with this CTA's first DEC_STAGES-1 KV tiles. Pages below the cache offset, the block
    // table and kv_start were written by EARLIER decode steps / the prefill, not by the kernel we are waiting for, and the
    // host-side offset (capture-time under a CUDA graph) is a lower bound of the real one: full tiles below it are immutable for
    // this launch. The KV stream then starts while the qkv projection drains instead of one memory round trip after the release
    // (in-kernel stamps: the attention window of a layer was 3.5 us longer than its stream). Only without split-KV, where the
    // first tile of the CTA does not depend on the real offset; otherwise the first tiles are at least pulled into L2.
    bool pre = false;
    if (p.past_host > 0 && p.pool && D == 96) {
        const int crow0 = b / p.row_div;
        const int kv00 = p.kv_start ? p.kv_start[crow0] : 0;
        const int tf = kv00 / 64, te = (p.past_host + 63) / 64, nt0 = max(te - tf, 0);
        if (p.early_fill && p.n_splits == 1 && tf + (DEC_STAGES - 1) <= p.past_host / 64) {
            pre = true;
            const int vz = p.zero * tid;                        // 0, but keeps the stage address in a vector register (see cp_async16_stream)
#pragma unroll 1
            for (int i = 0; i < DEC_STAGES - 1; i++) {
                issue_cached(i + vz, p.block_table[(size_t)crow0 * p.bt_stride + tf + i]);
                cp_async_commit();
            }
        } else {
            const int lo = tf + (int)(((long long)nt0 * split) / p.n_splits), hi = tf + (int)(((long long)nt0 * (split + 1)) / p.n_splits);
            const size_t he = (size_t)P3_PAGE * D, pe = 2 * (size_t)p.n_kv * he;
            for (int tix = lo; tix < min(hi, lo + DEC_STAGES - 1); tix++) {
                const int pg = p.block_table[(size_t)crow0 * p.bt_stride + tix];
                const uint8_t* kp = reinterpret_cast<const uint8_t*>(p.pool + (size_t)pg * pe + (size_t)kvh * he);
                const uint8_t* vp = kp + (size_t)p.n_kv * he * sizeof(bf16);
                for (int ln = tid; ln < (int)(he * sizeof(bf16) / 128); ln += 128) {
                    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(kp + (size_t)ln * 128));
                    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(vp + (size_t)ln * 128));
                }
            }
        }
    }
    pdl_wait();                                                 // q/k/v, past and the cache come from earlier kernels
    ATTN_TRACE_STAMP(p.trace, (int)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x), 1);
    const int past = p.past_dev ? *p.past_dev : p.past_host;
    const int crow = b / p.row_div;
    const int kv0 = p.kv_start ? p.kv_start[crow] : 0;
    const int s_total = past + p.L;
    const int32_t* bt = p.block_table + (size_t)crow * p.bt_stride;

    for (int idx = tid; idx < 16 * CPR; idx += 128) {
        int r = idx / CPR, c = idx % CPR;
        const bf16* src = p.q + ((size_t)b * p.L + (r < p.L ? r : 0)) * p.ldq + h * D + c * 8;
        cp_async16(sq + tile_off<D>(r, c), src, r < p.L ? 16 : 0);
    }
    // balanced split of the cached tiles [t_lo, t_hi); the last split also owns the "present" tile
    const int t_first = kv0 / 64, t_end = (past + 63) / 64;
    const int nt_all = max(t_end - t_first, 0);
    const int n_lo = t_first + (int)(((long long)nt_all * split) / p.n_splits);
    const int n_hi = t_first + (int)(((long long)nt_all * (split + 1)) / p.n_splits);
    const int n_cached = n_hi - n_lo;
    const bool has_present = (split == p.n_splits - 1);
    const int n_iter = n_cached + (has_present ? 1 : 0);

    auto issue_present = [&](int it) {
        const uint32_t sk = skv + (it % DEC_STAGES) * 2 * TILE, sv = sk + TILE;
        for (int idx = tid; idx < 16 * CPR; idx += 128) {
            int r = idx / CPR, c = idx % CPR;
            size_t tok = (size_t)b * p.L + (r < p.L ? r : 0);
            cp_async16(sk + tile_off<D>(r, c), p.k + tok * p.ldk + kvh * D + c * 8, r < p.L ? 16 : 0);
            cp_async16(sv + tile_off<D>(r, c), p.v + tok * p.ldv + kvh * D + c * 8, r < p.L ? 16 : 0);
        }
    };
    // page ids are fetched DEC_STAGES-1 tiles ahead of their use
    int page_next = (n_cached > 0) ? bt[n_lo] : 0;
    if (pre) {                                                  // the first DEC_STAGES-1 tiles are already in flight (n_cached >= DEC_STAGES-1)
        if (DEC_STAGES - 1 < n_cached) page_next = bt[n_lo + DEC_STAGES - 1];
        cp_async_commit();                                      // the Q tile
    } else {
#pragma unroll
        for (int i = 0; i < DEC_STAGES - 1; i++) {
            if (i < n_cached) {
                int pg = page_next;
                if (i + 1 < n_cached) page_next = bt[n_lo + i + 1];
                issue_cached(i, pg);
            } else if (i == n_cached && has_present) {
                issue_present(i);
            }
            cp_async_commit();                                  // group 0 also carries the Q tile
        }
    }

    uint32_t qf[D / 16][4];
    float o[D / 8][4];
#pragma unroll
    for (int dt = 0; dt < D / 8; dt++)
#pragma unroll
        for (int j = 0; j < 4; j++) o[dt][j] = 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    // hoisted ldmatrix lane offsets (row / swizzle parts that do not depend on the k-step)
    const int kr = warp * 16 + (lane >> 4) * 8 + (lane & 7);    // K row for QK^T
    const int vr = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;   // V row for PV
    const int pr_k = (lane >> 4) * 8 + (lane & 7), pr_v = (lane & 7) + ((lane >> 3) & 1) * 8;  // present tile rows

    // The o_proj weights (18.9 MB at Phi-3.5 sizes) are too small to stream efficiently on their own and
    // cannot be prefetched by their own kernel while this one holds the SMs' shared memory, so each
    // CTA pulls its slice of them into L2 three quarters of the way through its KV stream.
    const int pf_iter = (n_iter * 3) / 4;
    for (int it = 0; it < n_iter; it++) {
        if (pre && it == 0) cp_async_wait<0>();                 // early tiles + the Q group committed after them
        else cp_async_wait<DEC_STAGES - 2>();
        __syncthreads();
        if (it == pf_iter && p.l2_prefetch) {
            const int64_t n_cta = (int64_t)gridDim.x * gridDim.y * gridDim.z;
            const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            l2_prefetch_slice(p.l2_prefetch, p.l2_prefetch_bytes, cta, n_cta, tid, 128);
        }
        {   // refill the stage that was consumed in the previous iteration
            const int nx = it + DEC_STAGES - 1;
            if (nx < n_cached) {
                int pg = page_next;
                if (nx + 1 < n_cached) page_next = bt[n_lo + nx + 1];
                issue_cached(nx, pg);
            } else if (nx == n_cached && has_present) {
                issue_present(nx);
            }
            cp_async_commit();
        }
        if (it == 0) {
#pragma unroll
            for (int ks = 0; ks < D / 16; ks++) {
                int r = lane & 15, c = ks * 2 + (lane >> 4);
                ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], sq + tile_off<D>(r, c));
            }
        }
        const uint32_t sk = skv + (it % DEC_STAGES) * 2 * TILE, sv = sk + TILE;
        const bool present = it >= n_cached;
        if (present && warp != 0) continue;                     // 16 new keys: one warp
        // ---- S = Q K^T for this warp's 16 keys
        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const int krow = present ? pr_k : kr;
#pragma unroll
        for (int ks = 0; ks < D / 16; ks++) {
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(b0, b1, b2, b3, sk + tile_off<D>(krow, ks * 2 + ((lane >> 3) & 1)));
            mma_bf16_16816(s[0], qf[ks], b0, b1);
            mma_bf16_16816(s[1], qf[ks], b2, b3);
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
                    if (present) {
                        int qi = past + ((e & 2) ? g + 8 : g);
                        ok = (j >= kv0) && (j < s_total) && (j <= qi);   // left-pad keys of a short first call (past == 0)
                    } else {
                        ok = (j >= kv0) && (j < past);
                    }
                    if (!ok) s[nt][e] = -INFINITY;
                }
        }
        // ---- online softmax (rows g and g+8), exp2 domain, one FFMA + one MUFU per element
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
        if (c0 != 1.f || c1 != 1.f) {                           // warp-divergence free enough: skip when the max did not move
#pragma unroll
            for (int dt = 0; dt < D / 8; dt++) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
        }
        // ---- O += P V
        const uint32_t a[4] = {pack_bf16(s[0][0], s[0][1]), pack_bf16(s[0][2], s[0][3]),
                               pack_bf16(s[1][0], s[1][1]), pack_bf16(s[1][2], s[1][3])};
        const int vrow = present ? pr_v : vr;
#pragma unroll
        for (int dp = 0; dp < D / 16; dp++) {
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(b0, b1, b2, b3, sv + tile_off<D>(vrow, dp * 2 + (lane >> 4)));
            mma_bf16_16816(o[2 * dp], a, b0, b1);
            mma_bf16_16816(o[2 * dp + 1], a, b2, b3);
        }
    }
    cp_async_wait<0>();
    ATTN_TRACE_STAMP(p.trace, (int)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x), 3);
    __syncthreads();                                            // ring is free: reuse it for the warp merge

    float* sm_o = reinterpret_cast<float*>(smem + 16 * D * 2);  // [4][16][D]
    float* sm_m = sm_o + 4 * 16 * D;                            // [4][16]  (log2 domain, already scaled)
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
    if (g < p.L || g + 8 < p.L) {
#pragma unroll
        for (int dt = 0; dt < D / 8; dt++) {
            int d = dt * 8 + 2 * t;
            *reinterpret_cast<float2*>(sm_o + (warp * 16 + g) * D + d) = make_float2(o[dt][0], o[dt][1]);
            *reinterpret_cast<float2*>(sm_o + (warp * 16 + g + 8) * D + d) = make_float2(o[dt][2], o[dt][3]);
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
        // the last split of this (row, head) to arrive merges all partials (fixed order -> deterministic)
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
    ATTN_TRACE_STAMP(p.trace, (int)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x), 4);
}

// ------------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------------
static int fill_params(AttnParams& p, const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                       void* out, int64_t ldo, int B, int L, int n_heads, int n_kv, int hd, float scale, int causal,
                       int past, const int32_t* kv_start, const void* pool, const int32_t* block_table, int bt_stride,
                       int row_div) {
    P3_CHECK_ARG(hd == 96 || hd == 64, "attention: head_dim must be 96 or 64 (got %d)", hd);
    P3_CHECK_ARG(n_kv >= 1 && n_heads % n_kv == 0, "attention: n_heads must be a multiple of n_kv");
    P3_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "attention: strides must keep 16 B alignment");
    P3_CHECK_ARG(past == 0 || (pool && block_table), "attention: past > 0 needs a KV pool and block table");
    P3_CHECK_ARG(row_div >= 1, "attention: row_div must be >= 1");
    p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v;
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.out = (bf16*)out; p.ldo = ldo;
    p.B = B; p.L = L; p.n_heads = n_heads; p.n_kv = n_kv; p.hd = hd;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.causal = causal; p.past = past; p.past_host = past; p.past_dev = nullptr; p.kv_start = kv_start;
    p.pool = (const bf16*)pool; p.block_table = block_table; p.bt_stride = bt_stride; p.row_div = row_div;
    p.n_splits = 1; p.tiles_per_split = 0; p.ws_o = nullptr; p.ws_ml = nullptr; p.counters = nullptr;
    p.l2_prefetch = nullptr; p.l2_prefetch_bytes = 0; p.zero = 0;
    p.n_quant = 0; p.qcodes = nullptr; p.qmeta = nullptr; p.trace = nullptr; p.early_fill = 0;
    return 0;
}

bool attn_prefill_tc_eligible(const AttnParams& p);             // attention_tc.cu
int launch_prefill_tc(const AttnParams& p, cudaStream_t st);

extern "C" int p3_attention_prefill(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                    void* out, int64_t ldo, int B, int L, int n_heads, int n_kv, int hd, float scale,
                                    int causal, int past, const int32_t* kv_start, const void* pool,
                                    const int32_t* block_table, int bt_stride, int row_div, cudaStream_t st) {
    AttnParams p;
    if (fill_params(p, q, k, v, ldq, ldk, ldv, out, ldo, B, L, n_heads, n_kv, hd, scale, causal, past, kv_start, pool,
                    block_table, bt_stride, row_div)) return -1;
    if (B == 0 || L == 0) return 0;
    if (attn_prefill_tc_eligible(p)) return launch_prefill_tc(p, st);   // tcgen05 flash attention (attention_tc.cu)
    dim3 grid((L + 127) / 128, n_heads, B);
    int smem = (2 + 2 * PF_STAGES) * 64 * hd * 2;
    if (hd == 96) {
        static P3DevFlags flags; bool& set = flags.cur();
        if (!set) { cudaFuncSetAttribute(attn_prefill_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
        attn_prefill_kernel<96><<<grid, PF_THREADS, smem, st>>>(p);
    } else {
        static P3DevFlags flags; bool& set = flags.cur();
        if (!set) { cudaFuncSetAttribute(attn_prefill_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
        attn_prefill_kernel<64><<<grid, PF_THREADS, smem, st>>>(p);
    }
    P3_CHECK_LAUNCH("attention_prefill");
    return 0;
}

extern "C" int64_t p3_attention_decode_workspace(int B, int L, int n_heads, int hd, int n_splits) {
    (void)L;
    return (int64_t)B * n_heads * n_splits * 16 * (hd + 2) * 4 + (int64_t)B * n_heads * 4;   // partials + arrival counters
}

int launch_decode_q4_d96(AttnParams& p, cudaStream_t st);      // attention_q4.cu

template <int D>
static int launch_decode(AttnParams& p, cudaStream_t st) {
    dim3 grid(p.n_splits, p.n_heads, p.B);
    int smem = 16 * D * 2 + DEC_STAGES * 2 * 64 * D * 2;
    static P3DevFlags flags; bool& set = flags.cur();
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(attn_decode_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        P3_CHECK_ARG(e == cudaSuccess, "attention_decode: smem attribute: %s", cudaGetErrorString(e));
        set = true;
    }
#ifdef P3_TRACE_ATTN
    p.trace = p3_trace_slot();
#endif
    { static int v = -1; if (v < 0) { const char* e = getenv("P3_ATTN_EARLY"); v = e ? atoi(e) : 1; } p.early_fill = v; }
    p3_launch_pdl(attn_decode_kernel<D>, grid, dim3(128), (size_t)smem, st, p);
    P3_CHECK_LAUNCH("attention_decode");       // split partials are merged in-kernel by the last CTA to arrive
    return 0;
}

static int decode_common(AttnParams& p, int L, int past, int n_splits, void* workspace, bool q4, cudaStream_t st) {
    P3_CHECK_ARG(L >= 1 && L <= 16, "attention_decode: L must be in [1,16] (got %d)", L);
    P3_CHECK_ARG(n_splits >= 1 && (n_splits == 1 || workspace), "attention_decode: n_splits>1 needs a workspace");
    if (p.B == 0) return 0;
    int tiles_total = (past + 63) / 64;
    if (!p.past_dev && n_splits > tiles_total) n_splits = tiles_total > 0 ? tiles_total : 1;
    p.n_splits = n_splits;
    p.tiles_per_split = 0;                                      // balanced split is computed in the kernel
    p.ws_o = (float*)workspace;
    p.ws_ml = p.ws_o ? p.ws_o + (size_t)p.B * p.n_heads * n_splits * 16 * p.hd : nullptr;
    p.counters = p.ws_o ? reinterpret_cast<int*>(p.ws_ml + (size_t)p.B * p.n_heads * n_splits * 16 * 2) : nullptr;
    P3_CHECK_ARG(!q4 || p.hd == 96, "attention_decode_q4: only head_dim 96 is supported");
    if (p.hd == 96) return q4 ? launch_decode_q4_d96(p, st) : launch_decode<96>(p, st);   // q4: register-dequant kernel (attention_q4.cu)
    return launch_decode<64>(p, st);
}

extern "C" int p3_attention_decode(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                   void* out, int64_t ldo, int B, int L, int n_heads, int n_kv, int hd, float scale,
                                   int past, const int32_t* kv_start, const void* pool, const int32_t* block_table,
                                   int bt_stride, int row_div, int n_splits, void* workspace, const int32_t* past_dev,
                                   const void* l2_prefetch, int64_t l2_prefetch_bytes, cudaStream_t st) {
    AttnParams p;
    if (fill_params(p, q, k, v, ldq, ldk, ldv, out, ldo, B, L, n_heads, n_kv, hd, scale, 1, past, kv_start, pool,
                    block_table, bt_stride, row_div)) return -1;
    p.past_dev = past_dev;
    p.l2_prefetch = (const uint8_t*)l2_prefetch; p.l2_prefetch_bytes = l2_prefetch_bytes;
    return decode_common(p, L, past, n_splits, workspace, false, st);
}

extern "C" int p3_attention_decode_q4(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                      void* out, int64_t ldo, int B, int L, int n_heads, int n_kv, int hd, float scale,
                                      int past, int n_quant, const int32_t* kv_start, const void* pool, const void* qcodes,
                                      const void* qmeta, const int32_t* block_table, int bt_stride, int row_div,
                                      int n_splits, void* workspace, const int32_t* past_dev, const void* l2_prefetch,
                                      int64_t l2_prefetch_bytes, cudaStream_t st) {
    AttnParams p;
    if (fill_params(p, q, k, v, ldq, ldk, ldv, out, ldo, B, L, n_heads, n_kv, hd, scale, 1, past, kv_start, pool,
                    block_table, bt_stride, row_div)) return -1;
    P3_CHECK_ARG(n_quant % 64 == 0 && n_quant <= past, "attention_decode_q4: n_quant must be a multiple of 64 and <= past");
    P3_CHECK_ARG(n_quant == 0 || (qcodes && qmeta), "attention_decode_q4: q4 pools missing");
    p.n_quant = n_quant; p.qcodes = (const uint8_t*)qcodes; p.qmeta = (const bf16*)qmeta;
    p.past_dev = past_dev;
    p.l2_prefetch = (const uint8_t*)l2_prefetch; p.l2_prefetch_bytes = l2_prefetch_bytes;
    return decode_common(p, L, past, n_splits, workspace, true, st);
}
