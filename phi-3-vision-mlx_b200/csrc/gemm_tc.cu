// Dense bf16 GEMM on 5th-gen tensor cores: out[M,N] = X[M,K] . W[N,K]^T (+ fused epilogue).
// Replaces every prefill / ViT / projector nn.Linear and the patch-embed conv of the reference
// (phi.py:140-143,155-156,186-192,391,437-438,465-466,604).
//
// Structure (one CTA per SM, persistent over output tiles, 128 x BN tile, BK = 64):
//   warp 0   : TMA producer  — cp.async.bulk.tensor 2D, 128B swizzle, STAGES-deep mbarrier ring
//   warp 1   : MMA issuer    — one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                              (UMMA 128 x BN x 16), accumulators in TMEM (2 x BN columns, double buffered)
//   warp 2   : TMEM allocator
//   warps 4-11: epilogue     — tcgen05.ld 32x32b (two warps per TMEM lane quarter, half the columns each),
//                              fused bias / GELU / residual / SwiGLU / fp32 / row scatter
// Both operands are K-major (activations [M,K] and nn.Linear weights [N,K]), so no transposes.
#include "tc_common.cuh"
#include "../../include/phi3_b200.h"
#include <cstdlib>
#include <cstring>

struct GemmEpi {
    const bf16* bias;
    void* out;
    int64_t ldo;
    const void* resid;
    const int32_t* row_map;
    int kind;
    // fused RMSNorm (phi.py:478-479) of the INPUT rows: the gain is folded into W at load time, the per-row scale
    // rsqrt(sum_c ss_in[row][c] / K + eps) multiplies the accumulators; ss_in = sum-of-squares partials of the rows of X
    const float* ss_in; int n_ss_in; float eps;
    // RESIDUAL: sum of squares of every 32-column chunk written, ss_out[row][N/32] (feeds the next GEMM's ss_in)
    float* ss_out;
    // P3_EPI_ROPE_KV: SuRoPE (phi.py:418-423,487-507) + paged KV write (phi.py:542-548) on the qkv projection. W rows are
    // permuted per head so that every 32-column chunk holds 16 dims [16j,16j+16) and their rotary partners [half+16j, +16).
    const float *cosT, *sinT; int64_t tab_bstride;
    int L, n_heads, n_kv, hd, past, row_div, write_cache;
    const int32_t* past_dev;
    bf16* pool; const int32_t* block_table; int bt_stride;
    // split-K (small M: few output tiles, the GEMM is a weight stream and one SM ingests only ~1/148 of HBM bandwidth, so every
    // SM must stream): `ksplit` CTAs share an output tile, each accumulates a K slice, dumps fp32 partials to `ws`
    // [tile][ksplit][128][BN]; the last one to arrive (per-tile counter) sums them in slice order and runs the epilogue.
    int ksplit; float* ws; int* counters; int64_t ws_bytes;
    // outputs larger than the L2 can keep (batched prefill): store with the streaming (evict-first) policy so that the
    // write-allocated lines do not push the resident W band / X rows out of L2 (ncu: 2.7-3.1x DRAM over-read without it)
    int stream_out;
    int band_mb;                                           // rasterisation band (MB of W kept L2-resident); 0 = default
};
__device__ __forceinline__ void st_out16(void* p, const uint4& v, int stream) {
    if (stream) __stcs(reinterpret_cast<uint4*>(p), v); else *reinterpret_cast<uint4*>(p) = v;
}
#define P3_SPLITK_COUNTER_BYTES 4096
#define P3_EPI_ROPE_KV 8

__device__ __forceinline__ float epi_act(int kind, float x) {
    // CLIP fc1 and the projector run in fp32 in the reference (fp32 activations x bf16 weights
    // promote to fp32), so the activation is applied to the unrounded accumulator.
    if (kind == P3_EPI_QGELU) return __fdividef(x, 1.f + __expf(-1.702f * x));
    if (kind == P3_EPI_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    return x;
}

// Tile rasterisation: n-fastest inside bands of `nb` n-tiles so the band's W tiles (nb * BN * K * 2 B) stay
// L2-resident while A streams through once per band. (m-fastest order re-read A once per n-tile: ncu showed
// 2.1 GB of DRAM reads for a 157 MB problem.) Band size, ncu DRAM reads in the batched prefill (M = 16384):
// 48 MB bands -> qkv 727 MB, gate_up 733 MB; 32 MB -> 359 / 557 MB (a 48 MB band does not survive next to the X
// stream and the write-allocated output); for K = 8192 (4 MB per W tile, 74 concurrent tiles stream 72 MB per wave)
// nothing is retained across waves whatever the band, and the widest band that covers N (48 MB) reads least.
__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int nb, int& mt, int& nt) {
    const int band_tiles = nb * m_tiles;
    const int band = tile / band_tiles, rem = tile - band * band_tiles;
    const int nb_this = min(nb, n_tiles - band * nb);
    mt = rem / nb_this;
    nt = band * nb + rem - mt * nb_this;
}
__host__ __device__ __forceinline__ int band_width(int K, int BN, int n_tiles, int band_mb = 0) {
    long long per = (long long)BN * K * 2;
    int nb = (int)(((long long)(band_mb > 0 ? band_mb : (per <= (2ll << 20) ? 32 : 48)) << 20) / (per > 0 ? per : 1));
    return nb < 1 ? 1 : (nb > n_tiles ? n_tiles : nb);
}

template <int BN>
struct TcCfg {
    static constexpr int BM = 128, BK = 64;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
    // kind::f16 instruction descriptor: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), K-major both, N>>3 @17, M>>4 @24
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

// per-row RMSNorm scale from the producer's partial sums (fixed order: deterministic); 1 when the GEMM is not normed
__device__ __forceinline__ float epi_row_scale(const GemmEpi& ep, int row, int K, bool row_ok) {
    if (!ep.ss_in || !row_ok) return 1.f;
    const float* p = ep.ss_in + (size_t)row * ep.n_ss_in;
    float s = 0.f;
    if ((ep.n_ss_in & 3) == 0) {
        for (int c = 0; c < ep.n_ss_in; c += 4) { const float4 f = __ldcg(reinterpret_cast<const float4*>(p + c)); s += (f.x + f.y) + (f.z + f.w); }
    } else {
        for (int c = 0; c < ep.n_ss_in; c++) s += __ldcg(p + c);
    }
    return rsqrtf(s / (float)K + ep.eps);
}

// P3_EPI_ROPE_KV, per output row (token): everything that does not depend on the column chunk — computed once per tile,
// not once per 32-column chunk (the integer divisions by run-time L / page / hd cost more than the rotation itself)
struct RopeRow {
    bf16 *qrow, *kd, *vd;
    const float *cr, *sr;                                      // cos / sin rows of this token's position
};
__device__ __forceinline__ RopeRow epi_rope_row(const GemmEpi& ep, int row) {
    RopeRow r;
    const int hd = ep.hd, half = hd / 2;
    const int b = row / ep.L, pos = (ep.past_dev ? *ep.past_dev : ep.past) + (row - b * ep.L);
    const int crow = b / ep.row_div;
    r.qrow = reinterpret_cast<bf16*>(ep.out) + (size_t)row * ep.ldo;
    r.kd = r.vd = nullptr;
    if (ep.write_cache) {
        const int pg = pos / P3_PAGE;
        const int page = ep.block_table[(size_t)crow * ep.bt_stride + pg];
        r.kd = ep.pool + (size_t)page * kv_page_elems(ep.n_kv, hd) + (size_t)(pos - pg * P3_PAGE) * hd;
        r.vd = r.kd + (size_t)ep.n_kv * P3_PAGE * hd;
    }
    r.cr = ep.cosT + (size_t)crow * ep.tab_bstride + (size_t)pos * half;
    r.sr = ep.sinT + (size_t)crow * ep.tab_bstride + (size_t)pos * half;
    // the rows are read chunk by chunk after the accumulators arrive: have them in L1 by then (each miss is ~1 us of
    // exposed epilogue when a CTA has only one or two tiles)
    for (int o = 0; o < half * 4; o += 128) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(r.cr) + o));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(r.sr) + o));
    }
    return r;
}
// one 32-column chunk of the (row-permuted) qkv projection
__device__ __forceinline__ void epi_rope32(const GemmEpi& ep, const RopeRow& rr, int n0, float rs, const uint32_t* acc) {
    const int hd = ep.hd, half = hd / 2;
    const int head = n0 / hd, j = (n0 - head * hd) / 32;
    uint4 o1[2], o2[2];
    uint32_t* u1 = reinterpret_cast<uint32_t*>(o1);
    uint32_t* u2 = reinterpret_cast<uint32_t*>(o2);
    if (head < ep.n_heads + ep.n_kv) {                        // q or k head: rotate
        const float* cr = rr.cr + 16 * j;
        const float* sr = rr.sr + 16 * j;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(cr + i)), s4 = __ldg(reinterpret_cast<const float4*>(sr + i));
            const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
            float a1[4], a2[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {                     // the projection output is bf16 in the reference flow, rope runs in fp32
                const float x1 = bf16_round(__uint_as_float(acc[i + e]) * rs), x2 = bf16_round(__uint_as_float(acc[16 + i + e]) * rs);
                a1[e] = x1 * cs[e] - x2 * sn[e];
                a2[e] = x2 * cs[e] + x1 * sn[e];
            }
            u1[i / 2] = pack_bf16(a1[0], a1[1]); u1[i / 2 + 1] = pack_bf16(a1[2], a1[3]);
            u2[i / 2] = pack_bf16(a2[0], a2[1]); u2[i / 2 + 1] = pack_bf16(a2[2], a2[3]);
        }
        bf16* d1 = rr.qrow + head * hd + 16 * j;
        st_out16(d1, o1[0], ep.stream_out); st_out16(d1 + 8, o1[1], ep.stream_out);
        st_out16(d1 + half, o2[0], ep.stream_out); st_out16(d1 + half + 8, o2[1], ep.stream_out);
        if (rr.kd && head >= ep.n_heads) {
            bf16* k = rr.kd + (size_t)(head - ep.n_heads) * P3_PAGE * hd + 16 * j;
            reinterpret_cast<uint4*>(k)[0] = o1[0]; reinterpret_cast<uint4*>(k)[1] = o1[1];
            reinterpret_cast<uint4*>(k + half)[0] = o2[0]; reinterpret_cast<uint4*>(k + half)[1] = o2[1];
        }
    } else {                                                   // v head: 32 plain columns
#pragma unroll
        for (int i = 0; i < 8; i++) {
            u1[i] = pack_bf16(__uint_as_float(acc[2 * i]) * rs, __uint_as_float(acc[2 * i + 1]) * rs);
            u2[i] = pack_bf16(__uint_as_float(acc[16 + 2 * i]) * rs, __uint_as_float(acc[17 + 2 * i]) * rs);
        }
        bf16* d = rr.qrow + n0;
        st_out16(d, o1[0], ep.stream_out); st_out16(d + 8, o1[1], ep.stream_out);
        st_out16(d + 16, o2[0], ep.stream_out); st_out16(d + 24, o2[1], ep.stream_out);
        if (rr.vd) {
            bf16* v = rr.vd + (size_t)(head - ep.n_heads - ep.n_kv) * P3_PAGE * hd + 32 * j;
            reinterpret_cast<uint4*>(v)[0] = o1[0]; reinterpret_cast<uint4*>(v)[1] = o1[1];
            reinterpret_cast<uint4*>(v)[2] = o2[0]; reinterpret_cast<uint4*>(v)[3] = o2[1];
        }
    }
}

// epilogue for 32 consecutive accumulator columns of one output row (rs: RMSNorm row scale, 1 when not normed)
__device__ __forceinline__ void epi_store32(const GemmEpi& ep, int64_t orow, int n0, int N, const uint32_t* acc, float rs = 1.f) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(acc[i]) * rs;
    if (ep.bias) {
        if (n0 + 32 <= N) {
            uint4 bv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) bv[i] = __ldg(reinterpret_cast<const uint4*>(ep.bias + n0) + i);
            const uint32_t* bu = reinterpret_cast<const uint32_t*>(bv);
#pragma unroll
            for (int i = 0; i < 16; i++) { float2 f = unpack_bf16(bu[i]); v[2 * i] += f.x; v[2 * i + 1] += f.y; }
        } else {
            for (int i = 0; i < 32; i++) if (n0 + i < N) v[i] += __bfloat162float(ep.bias[n0 + i]);
        }
    }
    if (ep.kind == P3_EPI_F32 || ep.kind == P3_EPI_RESIDUAL_F32) {
        float* o = reinterpret_cast<float*>(ep.out) + orow * ep.ldo + n0;
        if (ep.kind == P3_EPI_RESIDUAL_F32) {
            const float* r = reinterpret_cast<const float*>(ep.resid) + orow * ep.ldo + n0;
            if (n0 + 32 <= N && (ep.ldo & 3) == 0) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float4 f = reinterpret_cast<const float4*>(r)[i];
                    v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w;
                }
            } else {
                for (int i = 0; i < 32; i++) if (n0 + i < N) v[i] += r[i];
            }
        }
        if (n0 + 32 <= N && (ep.ldo & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
            for (int i = 0; i < 32; i++) if (n0 + i < N) o[i] = v[i];
        }
        return;
    }
    bf16* o = reinterpret_cast<bf16*>(ep.out) + orow * ep.ldo + n0;
    const bool full = (n0 + 32 <= N) && ((ep.ldo & 7) == 0);
    if (ep.kind == P3_EPI_RESIDUAL) {
        const bf16* r = reinterpret_cast<const bf16*>(ep.resid) + orow * ep.ldo + n0;
        if (full) {
            uint4 rv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) rv[i] = reinterpret_cast<const uint4*>(r)[i];
            const uint32_t* ru = reinterpret_cast<const uint32_t*>(rv);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                float2 f = unpack_bf16(ru[i]);
                v[2 * i] = f.x + bf16_round(v[2 * i]);
                v[2 * i + 1] = f.y + bf16_round(v[2 * i + 1]);
            }
        } else {
            for (int i = 0; i < 32; i++) if (n0 + i < N) v[i] = __bfloat162float(r[i]) + bf16_round(v[i]);
        }
    } else if (ep.kind == P3_EPI_QGELU) {                     // one branch per chunk, not per element: with the kind test inside the
#pragma unroll                                                 // loop ptxas evaluates the erf polynomial for every element and selects
        for (int i = 0; i < 32; i++) v[i] = __fdividef(v[i], 1.f + __expf(-1.702f * v[i]));
    } else if (ep.kind == P3_EPI_GELU) {
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = 0.5f * v[i] * (1.f + erff(v[i] * 0.70710678118654752f));
    }
    if (full) {
        uint4 ov[4];
        uint32_t* ou = reinterpret_cast<uint32_t*>(ov);
#pragma unroll
        for (int i = 0; i < 16; i++) ou[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
#pragma unroll
        for (int i = 0; i < 4; i++) st_out16(o + 8 * i, ov[i], ep.stream_out);
        if (ep.ss_out) {                                       // sum of squares of the bf16 values just written
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < 16; i++) { const float2 f = unpack_bf16(ou[i]); sq += f.x * f.x + f.y * f.y; }
            ep.ss_out[(size_t)orow * (N / 32) + n0 / 32] = sq;
        }
    } else {
        float sq = 0.f;
        for (int i = 0; i < 32; i++) if (n0 + i < N) { o[i] = __float2bfloat16_rn(v[i]); const float f = bf16_round(v[i]); sq += f * f; }
        if (ep.ss_out) ep.ss_out[(size_t)orow * (N / 32) + n0 / 32] = sq;
    }
}

template <int BN>
__global__ void __launch_bounds__(384, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmEpi ep,
               int M, int N, int K) {
    using C = TcCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + C::STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + 2 + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
        smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + C::BM - 1) / C::BM, n_tiles = (N + BN - 1) / BN;
    const int ksplit = ep.ksplit > 1 ? ep.ksplit : 1;
    const int total = m_tiles * n_tiles * ksplit, kb_all = (K + C::BK - 1) / C::BK;   // work item = (tile, K slice)
    const int nb = band_width(K, BN, n_tiles, ep.band_mb);
    __shared__ int s_last;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // programmatic dependent launch: barrier init / TMEM allocation above overlap the previous kernel's tail; nothing
    // below (TMA loads of X, residual reads, output writes) may run before that kernel has completed
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x) {
                const int tile = item / ksplit, ks = item - tile * ksplit;
                int mt_, nt_;
                tile_coords(tile, m_tiles, n_tiles, nb, mt_, nt_);
                const int m_idx = mt_ * C::BM, n_idx = nt_ * BN;
                const int k0 = (int)((long long)kb_all * ks / ksplit), k1 = (int)((long long)kb_all * (ks + 1) / ksplit);
                for (int k = k0; k < k1; k++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    tma_load_2d(sa, &tmA, full_bar(stage), k * C::BK, m_idx);
                    tma_load_2d(sa + C::A_BYTES, &tmB, full_bar(stage), k * C::BK, n_idx);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // The whole warp stays converged and one elected lane issues: under `if (lane == 0)` nvcc keeps the UMMA
        // descriptors in vector registers and wraps every UTCHMMA in an ELECT / R2UR loop (~70 cycles per issue).
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            const int ks = item % ksplit;
            const int k0 = (int)((long long)kb_all * ks / ksplit), k1 = (int)((long long)kb_all * (ks + 1) / ksplit);
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int k = k0; k < k1; k++) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + C::A_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < C::BK / 16; kk++)   // +32 bytes (>>4 = 2) per UMMA_K inside the swizzle atom
                        tc_mma_bf16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, C::IDESC, ((k - k0) | kk) ? 1u : 0u);
                    tc_commit(empty_bar(stage));
                    if (k == k1 - 1) tc_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, half = (warp - 4) >> 2;       // TMEM lane quarter, column half
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            const int tile = item / ksplit, ks = item - tile * ksplit;
            int mt_, nt_;
            tile_coords(tile, m_tiles, n_tiles, nb, mt_, nt_);
            const int m_idx = mt_ * C::BM, n_idx = nt_ * BN;
            const int row = m_idx + q * 32 + lane;
            const bool row_ok = row < M;
            const int64_t orow = row_ok ? (ep.row_map ? (int64_t)ep.row_map[row] : (int64_t)row) : 0;
            const float rs = epi_row_scale(ep, row, K, row_ok);  // summed while the tile's MMAs run
            RopeRow rr;                                          // likewise: page lookup + cos/sin rows pulled into L1
            if (ep.kind == P3_EPI_ROPE_KV && row_ok) rr = epi_rope_row(ep, (int)orow);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            // chunk loader: 32 accumulator columns of this thread's row — straight from TMEM, or (split-K, last CTA of the
            // tile) the sum of all K slices' partials in slice order
            float* wsrow = ep.ws + ((size_t)tile * ksplit * C::BM + (size_t)(q * 32 + lane)) * BN;   // slice 0, this row
            bool from_ws = false;
            if (ksplit > 1) {
                float* mine = wsrow + (size_t)ks * C::BM * BN;
                for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(taddr + c0, v);                                     // (warp-collective: every lane loads)
                    if (!row_ok) continue;
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        reinterpret_cast<float4*>(mine + c0)[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                               __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                }
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));                                   // the accumulator stage is free again
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                __threadfence();
                asm volatile("bar.sync 2, 256;" ::: "memory");                  // the 8 epilogue warps
                if (threadIdx.x == 128) s_last = (atomicAdd(ep.counters + tile, 1) == ksplit - 1);
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (!s_last) continue;
                __threadfence();
                if (threadIdx.x == 128) ep.counters[tile] = 0;                  // ready for the next launch
                from_ws = true;
            }
            auto load32 = [&](int c, uint32_t* v) {
                if (!from_ws) { tc_ld32(taddr + c, v); return; }
                if (!row_ok) return;
                float a[32];
#pragma unroll
                for (int i = 0; i < 32; i++) a[i] = 0.f;
                for (int s2 = 0; s2 < ksplit; s2++) {
                    const float4* src = reinterpret_cast<const float4*>(wsrow + (size_t)s2 * C::BM * BN + c);
#pragma unroll
                    for (int i = 0; i < 8; i++) { const float4 f = __ldcg(src + i); a[4 * i] += f.x; a[4 * i + 1] += f.y; a[4 * i + 2] += f.z; a[4 * i + 3] += f.w; }
                }
#pragma unroll
                for (int i = 0; i < 32; i++) v[i] = __float_as_uint(a[i]);
            };
            if (ep.kind == P3_EPI_SWIGLU) {
                // interleaved weights: columns [0,BN/2) gate, [BN/2,BN) matching up
                for (int c0 = half * (BN / 4); c0 < (half + 1) * (BN / 4); c0 += 32) {
                    uint32_t g[32], u[32];
                    load32(c0, g);
                    load32(BN / 2 + c0, u);
                    const int on0 = n_idx / 2 + c0;
                    if (row_ok && on0 < N / 2) {
                        uint4 ov[4];
                        uint32_t* ou = reinterpret_cast<uint32_t*>(ov);
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            float g0 = bf16_round(__uint_as_float(g[2 * i]) * rs), g1 = bf16_round(__uint_as_float(g[2 * i + 1]) * rs);
                            float u0 = bf16_round(__uint_as_float(u[2 * i]) * rs), u1 = bf16_round(__uint_as_float(u[2 * i + 1]) * rs);
                            float a0 = bf16_round(g0 / (1.f + __expf(-g0))), a1 = bf16_round(g1 / (1.f + __expf(-g1)));
                            ou[i] = pack_bf16(a0 * u0, a1 * u1);
                        }
                        bf16* o = reinterpret_cast<bf16*>(ep.out) + orow * ep.ldo + on0;
                        if (on0 + 32 <= N / 2 && (ep.ldo & 7) == 0) {
#pragma unroll
                            for (int i = 0; i < 4; i++) st_out16(o + 8 * i, ov[i], ep.stream_out);
                        } else {
                            const bf16* ob = reinterpret_cast<const bf16*>(ov);
                            for (int i = 0; i < 32; i++) if (on0 + i < N / 2) o[i] = ob[i];
                        }
                    }
                }
            } else if (ep.kind == P3_EPI_ROPE_KV) {
                RopeRow rr;
                if (row_ok) rr = epi_rope_row(ep, (int)orow);
                for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
                    uint32_t v[32];
                    load32(c0, v);
                    if (row_ok && n_idx + c0 < N) epi_rope32(ep, rr, n_idx + c0, rs, v);
                }
            } else {
                for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
                    uint32_t v[32];
                    load32(c0, v);
                    if (row_ok && n_idx + c0 < N) epi_store32(ep, orow, n_idx + c0, N, v, rs);
                }
            }
            if (ksplit == 1) {
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// 2-CTA variant: a CTA pair (cluster 2x1, same TPC) computes a 256 x 256 tile with
// tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16). Each CTA TMA-loads its own 128 rows of A and only
// HALF of the W tile (128 rows); the tensor cores read the other half from the peer's shared
// memory, so L2->SM operand traffic per FLOP drops by a third (96 -> 64 B/cycle/SM at full rate) and
// the 6-stage ring holds 2x the K depth. The leader CTA (rank 0) issues every MMA; its commits are
// multicast to both CTAs' barriers; both CTAs' epilogue warps release the accumulator stage on the
// leader's barrier through the cluster shared window.
// ------------------------------------------------------------------------------------------
#define P3_PEER_MASK 0xFEFFFFFFu        // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cta0, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cta0), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_mc(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct Tc2Cfg {
    static constexpr int BM = 128, BN = 256, BK = 64, STAGES = 6;
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2;      // per CTA: own A rows + half of W rows
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmEpi ep,
                int M, int N, int K) {
    using C = Tc2Cfg;
    constexpr int BN = C::BN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + C::STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + 2 + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int m_pairs = (M + 2 * C::BM - 1) / (2 * C::BM), n_tiles = (N + BN - 1) / BN;
    const int total = m_pairs * n_tiles, kb = (K + C::BK - 1) / C::BK;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int nb = band_width(K, BN, n_tiles, ep.band_mb);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 512); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // programmatic dependent launch: barrier init / TMEM allocation above overlap the previous kernel's tail; nothing
    // below (TMA loads of X, residual reads, output writes) may run before that kernel has completed
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = cluster_id; tile < total; tile += n_clusters) {
                int mt_, nt_;
                tile_coords(tile, m_pairs, n_tiles, nb, mt_, nt_);
                const int m_idx = mt_ * (2 * C::BM) + (int)rank * C::BM;
                const int n_idx = nt_ * BN + (int)rank * (BN / 2);
                for (int k = 0; k < kb; k++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);   // both CTAs' bytes land on CTA 0's barrier
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t fb0 = full_bar(stage) & P3_PEER_MASK;
                    tma_load_2d_2sm(sa, &tmA, fb0, k * C::BK, m_idx);
                    tma_load_2d_2sm(sa + C::A_BYTES, &tmB, fb0, k * C::BK, n_idx);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {                                           // converged warp, elected issuing lane (see gemm_tc_kernel)
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < total; tile += n_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int k = 0; k < kb; k++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + C::A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < C::BK / 16; kk++)
                            tc_mma_bf16_2sm(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, C::IDESC, (k | kk) ? 1u : 0u);
                        tc_commit_2sm_mc(empty_bar(stage));                             // frees the slot in both CTAs
                        if (k == kb - 1) tc_commit_2sm_mc(tfull_bar(acc));              // accumulators ready in both CTAs
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, half = (warp - 4) >> 2;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = cluster_id; tile < total; tile += n_clusters) {
            int mt_, nt_;
            tile_coords(tile, m_pairs, n_tiles, nb, mt_, nt_);
            const int m_idx = mt_ * (2 * C::BM) + (int)rank * C::BM, n_idx = nt_ * BN;
            const int row = m_idx + q * 32 + lane;
            const bool row_ok = row < M;
            const int64_t orow = row_ok ? (ep.row_map ? (int64_t)ep.row_map[row] : (int64_t)row) : 0;
            const float rs = epi_row_scale(ep, row, K, row_ok);  // summed while the tile's MMAs run
            RopeRow rr;                                          // likewise: page lookup + cos/sin rows pulled into L1
            if (ep.kind == P3_EPI_ROPE_KV && row_ok) rr = epi_rope_row(ep, (int)orow);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            if (ep.kind == P3_EPI_SWIGLU) {
                for (int c0 = half * (BN / 4); c0 < (half + 1) * (BN / 4); c0 += 32) {
                    uint32_t g[32], u[32];
                    tc_ld32(taddr + c0, g);
                    tc_ld32(taddr + BN / 2 + c0, u);
                    const int on0 = n_idx / 2 + c0;
                    if (row_ok && on0 < N / 2) {
                        uint4 ov[4];
                        uint32_t* ou = reinterpret_cast<uint32_t*>(ov);
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            float g0 = bf16_round(__uint_as_float(g[2 * i]) * rs), g1 = bf16_round(__uint_as_float(g[2 * i + 1]) * rs);
                            float u0 = bf16_round(__uint_as_float(u[2 * i]) * rs), u1 = bf16_round(__uint_as_float(u[2 * i + 1]) * rs);
                            float a0 = bf16_round(g0 / (1.f + __expf(-g0))), a1 = bf16_round(g1 / (1.f + __expf(-g1)));
                            ou[i] = pack_bf16(a0 * u0, a1 * u1);
                        }
                        bf16* o = reinterpret_cast<bf16*>(ep.out) + orow * ep.ldo + on0;
                        if (on0 + 32 <= N / 2 && (ep.ldo & 7) == 0) {
#pragma unroll
                            for (int i = 0; i < 4; i++) st_out16(o + 8 * i, ov[i], ep.stream_out);
                        } else {
                            const bf16* ob = reinterpret_cast<const bf16*>(ov);
                            for (int i = 0; i < 32; i++) if (on0 + i < N / 2) o[i] = ob[i];
                        }
                    }
                }
            } else if (ep.kind == P3_EPI_ROPE_KV) {
                for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(taddr + c0, v);
                    if (row_ok && n_idx + c0 < N) epi_rope32(ep, rr, n_idx + c0, rs, v);
                }
            } else {
                for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(taddr + c0, v);
                    if (row_ok && n_idx + c0 < N) epi_store32(ep, orow, n_idx + c0, N, v, rs);
                }
            }
            tc_fence_before();
            mbar_arrive_cluster(tempty_bar(acc) & P3_PEER_MASK);                        // the leader's MMA warp owns this barrier
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// mma.sync cross-check GEMM (tests / bring-up only; not the product path)
// ------------------------------------------------------------------------------------------
#define FB_BM 64
#define FB_BN 64
#define FB_BK 32
__global__ void __launch_bounds__(128) gemm_mma_kernel(const bf16* __restrict__ X, int64_t ldx, const bf16* __restrict__ W,
                                                       int64_t ldw, GemmEpi ep, int M, int N, int K) {
    __shared__ __align__(16) bf16 sA[FB_BM][FB_BK + 8];
    __shared__ __align__(16) bf16 sB[FB_BN][FB_BK + 8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int m0 = blockIdx.y * FB_BM;
    const bool swiglu = ep.kind == P3_EPI_SWIGLU;
    // SwiGLU: this CTA produces 64 output columns; gate rows then (second pass) the matching up rows
    const int o0 = blockIdx.x * FB_BN;
    const int passes = swiglu ? 2 : 1;
    float acc[2][2][4][4];
    for (int pass = 0; pass < passes; pass++) {
        const int nrow0 = swiglu ? ((o0 / 128) * 256 + (o0 % 128) + pass * 128) : o0;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[pass][i][j][c] = 0.f;
        for (int k0 = 0; k0 < K; k0 += FB_BK) {
            for (int i = tid; i < FB_BM * FB_BK / 8; i += 128) {
                int r = i / (FB_BK / 8), c = (i % (FB_BK / 8)) * 8;
                uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
                if (m0 + r < M && k0 + c < K) va = *reinterpret_cast<const uint4*>(X + (size_t)(m0 + r) * ldx + k0 + c);
                if (nrow0 + r < N && k0 + c < K) vb = *reinterpret_cast<const uint4*>(W + (size_t)(nrow0 + r) * ldw + k0 + c);
                *reinterpret_cast<uint4*>(&sA[r][c]) = va;
                *reinterpret_cast<uint4*>(&sB[r][c]) = vb;
            }
            __syncthreads();
#pragma unroll
            for (int ks = 0; ks < FB_BK; ks += 16) {
                uint32_t a[2][4], b[4][2];
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    a[i][0] = *reinterpret_cast<const uint32_t*>(&sA[wm + i * 16 + g][ks + 2 * t]);
                    a[i][1] = *reinterpret_cast<const uint32_t*>(&sA[wm + i * 16 + g + 8][ks + 2 * t]);
                    a[i][2] = *reinterpret_cast<const uint32_t*>(&sA[wm + i * 16 + g][ks + 2 * t + 8]);
                    a[i][3] = *reinterpret_cast<const uint32_t*>(&sA[wm + i * 16 + g + 8][ks + 2 * t + 8]);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    b[j][0] = *reinterpret_cast<const uint32_t*>(&sB[wn + j * 8 + g][ks + 2 * t]);
                    b[j][1] = *reinterpret_cast<const uint32_t*>(&sB[wn + j * 8 + g][ks + 2 * t + 8]);
                }
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) mma_bf16_16816(acc[pass][i][j], a[i], b[j][0], b[j][1]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int row = m0 + wm + i * 16 + g + (c >> 1) * 8;
                int col = o0 + wn + j * 8 + 2 * t + (c & 1);
                if (row >= M) continue;
                int64_t orow = ep.row_map ? ep.row_map[row] : row;
                if (swiglu) {
                    if (col >= N / 2) continue;
                    float gb = bf16_round(acc[0][i][j][c]), ub = bf16_round(acc[1][i][j][c]);
                    float a = bf16_round(gb / (1.f + __expf(-gb)));
                    reinterpret_cast<bf16*>(ep.out)[orow * ep.ldo + col] = __float2bfloat16_rn(a * ub);
                    continue;
                }
                if (col >= N) continue;
                float v = acc[0][i][j][c] + (ep.bias ? __bfloat162float(ep.bias[col]) : 0.f);
                if (ep.kind == P3_EPI_F32) { reinterpret_cast<float*>(ep.out)[orow * ep.ldo + col] = v; continue; }
                if (ep.kind == P3_EPI_RESIDUAL_F32) {
                    reinterpret_cast<float*>(ep.out)[orow * ep.ldo + col] = v + reinterpret_cast<const float*>(ep.resid)[orow * ep.ldo + col];
                    continue;
                }
                if (ep.kind == P3_EPI_RESIDUAL) v = __bfloat162float(reinterpret_cast<const bf16*>(ep.resid)[orow * ep.ldo + col]) + bf16_round(v);
                else if (ep.kind == P3_EPI_QGELU || ep.kind == P3_EPI_GELU) v = epi_act(ep.kind, v);
                reinterpret_cast<bf16*>(ep.out)[orow * ep.ldo + col] = __float2bfloat16_rn(v);
            }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// Weight tensor maps are immutable for the life of a model: p3_gemm_plan_weights builds the three maps the kernels can ask
// for (box rows 256 / 128 for the single-CTA tiles, 128 = half a 256-wide tile for the CTA pair) ONCE into a caller-owned
// blob; p3_gemm_fused then skips cuTensorMapEncodeTiled for W (the X map still depends on the call's activation pointer).
struct GemmWPlan { uint64_t magic; const void* W; int64_t ldw; int32_t N, K; CUtensorMap t256, t128; };
static_assert(sizeof(GemmWPlan) <= P3_GEMM_WPLAN_BYTES, "p3_gemm weight plan blob too small");
#define P3_WPLAN_MAGIC 0x5033574D41505331ull

static int make_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows) {
    CUresult r = tc_encode_2d(tm, ptr, rows, K, ld, box_rows);
    P3_CHECK_ARG(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld ld=%lld", (int)r,
                 (long long)rows, (long long)K, (long long)ld);
    return 0;
}

static int num_sms() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
    return n;
}

static int cur_dev() { int d = 0; cudaGetDevice(&d); return d < 0 || d >= 64 ? 0 : d; }

template <int BN>
static int launch_tc(const void* X, int64_t ldx, const void* W, int64_t ldw, const GemmEpi& ep, int64_t M, int N, int K,
                     cudaStream_t st, const GemmWPlan* wp = nullptr) {
    using C = TcCfg<BN>;
    CUtensorMap ta, tb;
    if (make_tmap(&ta, X, M, K, ldx, C::BM)) return -1;
    if (wp && BN >= 128) memcpy(&tb, (BN == 256) ? &wp->t256 : &wp->t128, sizeof(tb));
    else if (make_tmap(&tb, W, N, K, ldw, BN)) return -1;      // 64-row boxes are not part of the plan blob: encoded per call
    static bool attr_set[64] = {};                             // the attribute is per device
    if (!attr_set[cur_dev()]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        P3_CHECK_ARG(e == cudaSuccess, "gemm: cannot set %d B dynamic smem: %s", C::SMEM, cudaGetErrorString(e));
        attr_set[cur_dev()] = true;
    }
    int64_t tiles = ((M + C::BM - 1) / C::BM) * ((N + BN - 1) / BN);
    GemmEpi e2 = ep;
    e2.ksplit = 1;
    if (ep.ws && tiles * 2 < num_sms() && (int64_t)BN * K * 2 > (7ll << 18) * BN / 128) {
        // few output tiles AND a long K (each CTA would stream > 1.75 MB of W alone at ~45 KB/us, e.g. down_proj at M <= 320:
        // 24 tiles x 2 MB): split K so that tiles * ksplit ~ #SM (>= 4 k-blocks each). Measured (tools/gemm_smallm.py): 43 -> 30 us
        // at M = 80, 40 -> 21 us at M = 16; for K = 3072 shapes the partial round trip costs more than it saves, so they stay unsplit.
        const int kb = (K + C::BK - 1) / C::BK;
        int ks = (int)(num_sms() / tiles);
        if (ks > 8) ks = 8;
        if (ks > kb / 4) ks = kb / 4;
        const int64_t need = P3_SPLITK_COUNTER_BYTES + tiles * ks * (int64_t)C::BM * BN * 4;
        if (ks >= 2 && tiles * 4 <= P3_SPLITK_COUNTER_BYTES && need <= ep.ws_bytes) {
            e2.ksplit = ks;
            e2.counters = reinterpret_cast<int*>(ep.ws);
            e2.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ep.ws) + P3_SPLITK_COUNTER_BYTES);
        }
    }
    const int64_t items = tiles * e2.ksplit;
    unsigned grid = (unsigned)(items < num_sms() ? items : num_sms());
    p3_launch_pdl(gemm_tc_kernel<BN>, dim3(grid), dim3(384), (size_t)C::SMEM, st, ta, tb, e2, (int)M, N, K);
    P3_CHECK_LAUNCH("gemm_tc");
    return 0;
}

static int gemm_2cta_mode() {                  // P3_GEMM_2CTA=0 forces the single-CTA kernel (A/B testing)
    static int v = -1;
    if (v < 0) { const char* e = getenv("P3_GEMM_2CTA"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}

static int launch_tc2(const void* X, int64_t ldx, const void* W, int64_t ldw, const GemmEpi& ep, int64_t M, int N, int K,
                      cudaStream_t st, const GemmWPlan* wp = nullptr) {
    using C = Tc2Cfg;
    CUtensorMap ta, tb;
    if (make_tmap(&ta, X, M, K, ldx, C::BM)) return -1;
    if (wp) memcpy(&tb, &wp->t128, sizeof(tb));               // half of a 256-wide W tile per CTA
    else if (make_tmap(&tb, W, N, K, ldw, C::BN / 2)) return -1;
    static bool attr_set[64] = {};
    if (!attr_set[cur_dev()]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        P3_CHECK_ARG(e == cudaSuccess, "gemm: cannot set %d B dynamic smem: %s", C::SMEM, cudaGetErrorString(e));
        attr_set[cur_dev()] = true;
    }
    int64_t tiles = ((M + 2 * C::BM - 1) / (2 * C::BM)) * ((N + C::BN - 1) / C::BN);
    int64_t clusters = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
    p3_launch_pdl(gemm_tc2_kernel, dim3((unsigned)(2 * clusters)), dim3(384), (size_t)C::SMEM, st, ta, tb, ep, (int)M, N, K);
    P3_CHECK_LAUNCH("gemm_tc2");
    return 0;
}

static int gemm_dispatch(const void* X, int64_t ldx, const void* W, int64_t ldw, const GemmEpi& ep, int64_t M, int N, int K,
                         int impl, cudaStream_t st, const GemmWPlan* wp) {
    if (impl == 1) {
        P3_CHECK_ARG(ep.kind != P3_EPI_ROPE_KV && !ep.ss_in && !ep.ss_out, "gemm: the mma.sync cross-check kernel has no fused norm / rope epilogue");
        int ncols = ep.kind == P3_EPI_SWIGLU ? N / 2 : N;
        dim3 grid((ncols + FB_BN - 1) / FB_BN, (unsigned)((M + FB_BM - 1) / FB_BM));
        gemm_mma_kernel<<<grid, 128, 0, st>>>((const bf16*)X, ldx, (const bf16*)W, ldw, ep, (int)M, N, K);
        P3_CHECK_LAUNCH("gemm_mma");
        return 0;
    }
    int64_t m_tiles = (M + 127) / 128;
    GemmEpi eps = ep;
    {
        static int mode = -1;                              // P3_GEMM_STREAM=0/1 forces the policy off/on (A/B only)
        if (mode < 0) { const char* e = getenv("P3_GEMM_STREAM"); mode = e ? (e[0] == '0' ? 0 : 1) : 2; }
        const int64_t out_bytes = M * (int64_t)(ep.kind == P3_EPI_SWIGLU ? N / 2 : N) * 2;
        eps.stream_out = mode == 2 ? (out_bytes >= (64ll << 20)) : mode;
        static int band = -1;                              // P3_GEMM_BAND_MB (A/B only)
        if (band < 0) { const char* e = getenv("P3_GEMM_BAND_MB"); band = e ? atoi(e) : 0; }
        eps.band_mb = band;
    }
    if (impl == 0 && gemm_2cta_mode() && M > 128 && ((m_tiles + 1) / 2) * ((N + 255) / 256) >= num_sms() / 4)
        return launch_tc2(X, ldx, W, ldw, eps, M, N, K, st, wp);
    bool small = m_tiles * ((N + 255) / 256) < 2 * num_sms();
    if (ep.kind == P3_EPI_SWIGLU || !small) return launch_tc<256>(X, ldx, W, ldw, eps, M, N, K, st, wp);
    // one row tile and fewer 128-wide column tiles than SMs (M <= 128 against the 3072..9216-row matrices: constrain steps): the GEMM
    // is a weight stream and a CTA only pulls what its TMA ring turns over per round trip -> 64-row W tiles put twice as many CTAs
    // on it (tools/gemm_smallm.py, M = 80: qkv 18.9 -> 16.8 us, o_proj 19.3 -> 17.2, down_proj with split-K 28.6 -> 21.8). Not for
    // more row tiles: every CTA re-reads its X rows over the whole K, which outweighs the gain (M = 320 down_proj 36.6 -> 39.4 us).
    static int bn64 = -1;
    if (bn64 < 0) { const char* e = getenv("P3_GEMM_BN64"); bn64 = e ? atoi(e) : 1; }
    if (bn64 && m_tiles == 1 && (N + 127) / 128 < num_sms()) return launch_tc<64>(X, ldx, W, ldw, eps, M, N, K, st, wp);
    return launch_tc<128>(X, ldx, W, ldw, eps, M, N, K, st, wp);
}

extern "C" int p3_gemm(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* out, int64_t ldo,
                       const void* resid, const int32_t* row_map, int64_t M, int N, int K, int epi, int impl,
                       cudaStream_t st) {
    P3_CHECK_ARG(epi >= P3_EPI_NONE && epi <= P3_EPI_RESIDUAL_F32, "gemm: unknown epilogue %d", epi);
    P3_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0, "gemm: ldx/ldw must be multiples of 8 elements (16 B)");
    P3_CHECK_ARG(((uintptr_t)X & 15) == 0 && ((uintptr_t)W & 15) == 0, "gemm: X/W must be 16-byte aligned");
    P3_CHECK_ARG((epi != P3_EPI_RESIDUAL && epi != P3_EPI_RESIDUAL_F32) || resid, "gemm: residual epilogue needs resid");
    P3_CHECK_ARG(epi != P3_EPI_SWIGLU || (N % 256 == 0 && !bias), "gemm: SwiGLU needs N %% 256 == 0 and no bias");
    P3_CHECK_ARG(M < (1ll << 31), "gemm: M too large");
    if (M == 0) return 0;
    GemmEpi ep{};
    ep.bias = (const bf16*)bias; ep.out = out; ep.ldo = ldo; ep.resid = resid; ep.row_map = row_map; ep.kind = epi;
    return gemm_dispatch(X, ldx, W, ldw, ep, M, N, K, impl, st, nullptr);
}

extern "C" int p3_gemm_plan_weights(const void* W, int64_t ldw, int N, int K, void* plan) {
    P3_CHECK_ARG(W && plan && ldw % 8 == 0 && ((uintptr_t)W & 15) == 0, "gemm_plan_weights: bad arguments");
    P3_CHECK_ARG(((uintptr_t)plan & 63) == 0, "gemm_plan_weights: the plan blob must be 64-byte aligned");
    GemmWPlan* p = reinterpret_cast<GemmWPlan*>(plan);
    p->W = W; p->ldw = ldw; p->N = N; p->K = K;
    if (make_tmap(&p->t256, W, N, K, ldw, 256)) return -1;
    if (make_tmap(&p->t128, W, N, K, ldw, 128)) return -1;
    p->magic = P3_WPLAN_MAGIC;
    return 0;
}

// struct-argument entry (SURVEY.md par. 8b): everything p3_gemm does + the fused prefill epilogues
extern "C" int p3_gemm_fused(const p3_gemm_args* a, cudaStream_t st) {
    P3_CHECK_ARG(a, "gemm_fused: null args");
    const int epi = a->epi;
    P3_CHECK_ARG((epi >= P3_EPI_NONE && epi <= P3_EPI_RESIDUAL_F32) || epi == P3_EPI_ROPE_KV, "gemm_fused: unknown epilogue %d", epi);
    P3_CHECK_ARG(a->ldx % 8 == 0 && a->ldw % 8 == 0, "gemm_fused: ldx/ldw must be multiples of 8 elements (16 B)");
    P3_CHECK_ARG(((uintptr_t)a->X & 15) == 0 && ((uintptr_t)a->W & 15) == 0, "gemm_fused: X/W must be 16-byte aligned");
    P3_CHECK_ARG((epi != P3_EPI_RESIDUAL && epi != P3_EPI_RESIDUAL_F32) || a->resid, "gemm_fused: residual epilogue needs resid");
    P3_CHECK_ARG(epi != P3_EPI_SWIGLU || (a->N % 256 == 0 && !a->bias), "gemm_fused: SwiGLU needs N %% 256 == 0 and no bias");
    P3_CHECK_ARG(!a->ss_out || (epi == P3_EPI_RESIDUAL && a->N % 32 == 0 && !a->row_map), "gemm_fused: ss_out needs the bf16 residual epilogue, N %% 32 == 0");
    P3_CHECK_ARG(!a->ss_in || (a->n_ss_in >= 1 && !a->row_map), "gemm_fused: ss_in needs n_ss_in >= 1 and no row map");
    P3_CHECK_ARG(a->M < (1ll << 31), "gemm_fused: M too large");
    if (a->M == 0) return 0;
    GemmEpi ep{};
    ep.bias = (const bf16*)a->bias; ep.out = a->out; ep.ldo = a->ldo; ep.resid = a->resid; ep.row_map = a->row_map; ep.kind = epi;
    ep.ss_in = a->ss_in; ep.n_ss_in = a->n_ss_in; ep.eps = a->eps; ep.ss_out = a->ss_out;
    if (epi == P3_EPI_ROPE_KV) {
        P3_CHECK_ARG(a->hd % 32 == 0 && a->N == (a->n_heads + 2 * a->n_kv) * a->hd && a->ldo % 8 == 0 && !a->bias && !a->row_map,
                     "gemm_fused: ROPE_KV needs head_dim %% 32 == 0, N = (n_heads + 2 n_kv) * head_dim, no bias / row map");
        P3_CHECK_ARG(a->cosT && a->sinT && a->L >= 1 && a->row_div >= 1 && a->M % a->L == 0, "gemm_fused: ROPE_KV needs rope tables and M = B * L");
        P3_CHECK_ARG(!a->write_cache || (a->pool && a->block_table), "gemm_fused: cache write needs pool and block table");
        ep.cosT = a->cosT; ep.sinT = a->sinT; ep.tab_bstride = a->tab_bstride; ep.L = a->L; ep.n_heads = a->n_heads; ep.n_kv = a->n_kv;
        ep.hd = a->hd; ep.past = a->past; ep.row_div = a->row_div; ep.write_cache = a->write_cache; ep.past_dev = a->past_dev;
        ep.pool = (bf16*)a->pool; ep.block_table = a->block_table; ep.bt_stride = a->bt_stride;
    }
    ep.ws = reinterpret_cast<float*>(a->splitk_ws); ep.ws_bytes = a->splitk_ws ? a->splitk_ws_bytes : 0;   // carved up by launch_tc
    const GemmWPlan* wp = reinterpret_cast<const GemmWPlan*>(a->w_plan);
    if (wp) P3_CHECK_ARG(wp->magic == P3_WPLAN_MAGIC && wp->W == a->W && wp->ldw == a->ldw && wp->N == a->N && wp->K == a->K,
                         "gemm_fused: weight plan does not describe this W");
    return gemm_dispatch(a->X, a->ldx, a->W, a->ldw, ep, a->M, a->N, a->K, a->impl, st, wp);
}
