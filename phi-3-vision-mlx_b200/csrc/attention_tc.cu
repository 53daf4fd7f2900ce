// Flash attention for many query rows (prefill, ViT) on the 5th-gen tensor cores.
// Same contract as attn_prefill_kernel (attention.cu): softmax(QK^T * scale + mask) V with the
// causal + left-pad predicate of phi.py:550-563, keys [0,past) from the paged pool and keys
// [past,past+L) from the fresh qkv rows (phi.py:454-457; ViT: phi.py:148 with causal = 0).
//
// One CTA = 128 query rows of one (sequence, head); key tiles of 128.
//   warps 0-7 : softmax — two threads per query row (TMEM lane = row; warps w and w+4 take 64 keys each and
//               exchange the tile maximum through shared memory): tcgen05.ld of the S row into registers, running max with
//               lazy rescaling (O and l are only rescaled when the max grows by more than 2^8), exp2 via one
//               FFMA + one MUFU, P rounded to bf16 and written to shared memory in the K-major
//               128B-swizzled UMMA layout (double buffered); final O / l and the store.
//   warp 8    : TMA producer — Q once, then K and V tiles through two independent mbarrier rings
//               (K is released as soon as S = QK^T retires, V after O += PV). Pool pages are 64-key boxes.
//   warp 9    : TMEM allocation + MMA issuer — S(j+1) = Q K(j+1)^T is issued before O += P(j) V(j) so the
//               tensor pipe computes the next scores while the softmax warps work on the current ones.
// TMEM: S double buffered (2 x 128 columns), O (D columns). Operands: Q, K as K-major SW128 tiles (one or
// two 64-dim boxes; head_dim 96 uses half of the second box), V consumed in place as an MN-major B operand
// (its [key][dim] layout is already N-contiguous), P as a K-major A operand.
#include "attn_common.cuh"
#include "tc_common.cuh"
#include "../../include/phi3_b200.h"
#include <cstdlib>

template <int D>
struct FaCfg {
    static constexpr int NB = (D + 63) / 64;            // 64-dim boxes per tile row
    static constexpr int BOX = 128 * 128;               // bytes: 128 rows x 64 bf16, 128B swizzle
    static constexpr int Q_BYTES = NB * BOX, KV_BYTES = NB * BOX, P_BYTES = 2 * BOX;   // P is double buffered
    static constexpr int STAGES = (D == 96) ? 2 : 3;
    static constexpr int XCH_BYTES = (D == 96) ? 0 : 2048;   // D = 96 keeps the exchange slots in unused Q columns
    static constexpr int SMEM = Q_BYTES + 2 * STAGES * KV_BYTES + 2 * P_BYTES + 1024 + 256 + XCH_BYTES;
    static constexpr int TMEM_COLS = 512, S_COL = 0, O_COL = 256;
    static constexpr uint32_t IDESC_QK = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    static constexpr uint32_t IDESC_PV = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((128u >> 4) << 24);
};

#define FA_THREADS 320
#define FA_W_TMA 8
#define FA_W_MMA 9
#define FA_RESCALE_LOG2 8.0f

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmPool, AttnParams p) {
    using C = FaCfg<D>;
    constexpr int ST = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base, sK0 = sQ + C::Q_BYTES, sV0 = sK0 + ST * C::KV_BYTES, sP = sV0 + ST * C::KV_BYTES;
    const uint32_t bars = sP + 2 * C::P_BYTES;
    const uint32_t q_full = bars;
    auto k_full = [&](int s) { return bars + 8u * (1 + s); };
    auto k_empty = [&](int s) { return bars + 8u * (1 + ST + s); };
    auto v_full = [&](int s) { return bars + 8u * (1 + 2 * ST + s); };
    auto v_empty = [&](int s) { return bars + 8u * (1 + 3 * ST + s); };
    auto s_full = [&](int s) { return bars + 8u * (1 + 4 * ST + s); };
    auto s_free = [&](int s) { return bars + 8u * (3 + 4 * ST + s); };
    auto p_full = [&](int j) { return bars + 8u * (10 + 4 * ST + (j & 1)); };   // P(j) written (alternating, like o_done)
    // O += P(j) V(j) retired; two barriers alternating with the tile parity so that a waiter that lags by up to
    // two tiles never sees an aliased phase parity
    auto o_done = [&](int j) { return bars + 8u * (6 + 4 * ST + (j & 1)); };
    const uint32_t tmem_slot = bars + 8u * (8 + 4 * ST);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = (gridDim.x - 1 - blockIdx.x) * 128, h = blockIdx.y, b = blockIdx.z;   // heavy (late) row blocks first
    const int kvh = h / (p.n_heads / p.n_kv);
    const int past = p.past, s_total = past + p.L;
    const int crow = b / p.row_div;
    const int kv0 = p.kv_start ? p.kv_start[crow] : 0;
    const int n_begin = kv0 / 128;
    int n_end = (s_total + 127) / 128;
    if (p.causal) n_end = min(n_end, (past + min(i0 + 127, p.L - 1)) / 128 + 1);
    const int n_tiles = max(n_end - n_begin, 0);

    if (warp == FA_W_TMA && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmPool)) : "memory");
    }
    if (warp == FA_W_MMA) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int s = 0; s < ST; s++) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
            for (int s = 0; s < 2; s++) { mbar_init(s_full(s), 1); mbar_init(s_free(s), 8); }
            mbar_init(p_full(0), 8); mbar_init(p_full(1), 8);    // one elected arrival per softmax warp
            mbar_init(o_done(0), 1); mbar_init(o_done(1), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == FA_W_TMA) {
        // ---------------- TMA producer ----------------
        if (lane == 0 && n_tiles > 0) {
            mbar_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
            for (int bx = 0; bx < C::NB; bx++) tma_load_2d(sQ + bx * C::BOX, &tmQ, q_full, h * D + 64 * bx, b * p.L + i0);
            const int32_t* bt = p.block_table ? p.block_table + (size_t)crow * p.bt_stride : nullptr;
            auto load_tile = [&](int n, int kv, uint32_t dst, uint32_t bar) {
                mbar_expect_tx(bar, C::KV_BYTES);
                if (n * 128 < past) {                            // two 64-key pages of the pool
#pragma unroll
                    for (int pg = 0; pg < 2; pg++) {
                        const int row = ((bt[2 * n + pg] * 2 + kv) * p.n_kv + kvh) * P3_PAGE;
#pragma unroll
                        for (int bx = 0; bx < C::NB; bx++)
                            tma_load_2d(dst + bx * C::BOX + pg * (64 * 128), &tmPool, bar, 64 * bx, row);
                    }
                } else {                                         // 128 fresh rows of the qkv buffer
                    const int tok = b * p.L + (n * 128 - past);
#pragma unroll
                    for (int bx = 0; bx < C::NB; bx++)
                        tma_load_2d(dst + bx * C::BOX, kv ? &tmV : &tmK, bar, kvh * D + 64 * bx, tok);
                }
            };
            for (int it = 0; it < n_tiles; it++) {
                const int st = it % ST;
                const uint32_t ph = (uint32_t)(it / ST) & 1u;
                mbar_wait(k_empty(st), ph ^ 1u);
                load_tile(n_begin + it, 0, sK0 + st * C::KV_BYTES, k_full(st));
                mbar_wait(v_empty(st), ph ^ 1u);
                load_tile(n_begin + it, 1, sV0 + st * C::KV_BYTES, v_full(st));
            }
        }
    } else if (warp == FA_W_MMA) {
        // ---------------- MMA issuer ----------------
        if (lane == 0 && n_tiles > 0) {
            mbar_wait(q_full, 0);
            for (int it = 0; it <= n_tiles; it++) {
                if (it < n_tiles) {                              // S(it) = Q K(it)^T
                    const int st = it % ST, sb = it & 1;
                    mbar_wait(k_full(st), (uint32_t)(it / ST) & 1u);
                    mbar_wait(s_free(sb), ((uint32_t)(it >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + C::S_COL + sb * 128;
#pragma unroll
                    for (int kk = 0; kk < D / 16; kk++) {
                        const uint64_t ad = umma_desc_sw128(sQ + (kk >> 2) * C::BOX) + 2 * (kk & 3);
                        const uint64_t bd = umma_desc_sw128(sK0 + st * C::KV_BYTES + (kk >> 2) * C::BOX) + 2 * (kk & 3);
                        tc_mma_bf16(d_tmem, ad, bd, C::IDESC_QK, kk ? 1u : 0u);
                    }
                    tc_commit(k_empty(st));
                    tc_commit(s_full(sb));
                }
                if (it > 0) {                                    // O += P(j) V(j)
                    const int j = it - 1, st = j % ST;
                    mbar_wait(v_full(st), (uint32_t)(j / ST) & 1u);
                    mbar_wait(p_full(j), (uint32_t)(j >> 1) & 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + C::O_COL;
#pragma unroll
                    for (int kk = 0; kk < 8; kk++) {             // 16 keys per UMMA
                        const uint64_t ad = umma_desc_sw128(sP + (j & 1) * C::P_BYTES + (kk >> 2) * C::BOX) + 2 * (kk & 3);
                        const uint64_t bd = umma_desc_mn_sw128(sV0 + st * C::KV_BYTES + kk * (16 * 128), C::BOX, 1024);
                        tc_mma_bf16(d_tmem, ad, bd, C::IDESC_PV, (j | kk) ? 1u : 0u);
                    }
                    tc_commit(v_empty(st));
                    tc_commit(o_done(j));
                }
            }
        }
    } else {
        // ---------------- softmax: two threads per query row (warps w and w+4 split the 128 keys) ----------------
        const int wq = warp & 3, half = warp >> 2;
        const int row = wq * 32 + lane;
        const int qi = past + i0 + row;
        const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
        const float sl = p.scale_log2;
        float m = -INFINITY, l = 0.f;
        const int rx = row & 7;
        const uint32_t p_row = sP + half * C::BOX + row * 128;   // this thread's 64 keys are one 64-column box of P
        // 16 B exchange slot per row: float [parity][half]. D = 96: logical chunk 4 of Q box 1 (dims 96..103 of the
        // padded tile, written once by the Q load and never read by an MMA); D = 64: a separate region.
        const uint32_t xch = (D == 96) ? (sQ + C::BOX + row * 128 + ((4 ^ rx) << 4)) : (bars + 256 + row * 16);
        auto exchange = [&](int par, float mine) -> float {      // returns the partner thread's value
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch + par * 8 + half * 4), "f"(mine) : "memory");
            asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
            float other;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xch + par * 8 + (half ^ 1) * 4) : "memory");
            return other;
        };
        for (int it = 0; it < n_tiles; it++) {
            const int sb = it & 1, j0 = (n_begin + it) * 128 + half * 64;
            mbar_wait(s_full(sb), (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_s = t_lane + C::S_COL + sb * 128 + half * 64;
            // this thread's half of the S row (64 fp32) lives in registers: one TMEM round trip per tile
            uint32_t s[64];
            tc_ld32_nowait(t_s, s);
            tc_ld32_nowait(t_s + 32, s + 32);
            tc_wait_ld();
            tc_reg_fence32(s);
            tc_reg_fence32(s + 32);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free(sb));              // S(it+2) may overwrite the buffer now
            const bool need_mask = (j0 < kv0) || (j0 + 63 >= s_total) || (p.causal && j0 + 63 > past + i0 + wq * 32);
            if (need_mask) {
#pragma unroll
                for (int i = 0; i < 64; i++) {
                    const int j = j0 + i;
                    const bool ok = (j >= kv0) && (j < s_total) && (!p.causal || j <= qi);
                    if (!ok) s[i] = 0xff800000u;
                }
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(s[i])); mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(s[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
            }
            float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            mx = fmaxf(mx, exchange(it & 1, mx));                // both threads of a row now hold the same tile maximum
            const bool grow = mx * sl > m * sl + FA_RESCALE_LOG2;  // false when both are -inf
            const bool resc = grow && it > 0 && m != -INFINITY;
            float corr = 1.f;
            if (grow) {
                corr = (m == -INFINITY) ? 0.f : ex2_approx((m - mx) * sl);
                m = mx;
                l *= corr;
            }
            if (it > 1) mbar_wait(o_done(it - 2), (uint32_t)((it - 2) >> 1) & 1u);   // PV(it-2) retired: P buffer (it & 1) is free again
            if (__any_sync(0xffffffffu, resc)) {                 // rare: the running max grew by more than 2^8
                mbar_wait(o_done(it - 1), (uint32_t)((it - 1) >> 1) & 1u);   // O is complete up to tile it-1 and PV(it) waits for our P
                tc_fence_after();
                const float cf = resc ? corr : 1.f;
#pragma unroll 1
                for (int c = half; c < D / 32; c += 2) {         // the two threads of a row split the O columns
                    uint32_t v[32];
                    tc_ld32(t_lane + C::O_COL + 32 * c, v);
#pragma unroll
                    for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * cf);
                    tc_st32(t_lane + C::O_COL + 32 * c, v);
                }
                tc_fence_before();
            }
            // P = exp2(S * scale - m * scale) -> bf16 -> shared memory (K-major SW128 A operand)
            const float mu = (m == -INFINITY) ? 0.f : m * sl;
            float ls0 = 0.f, ls1 = 0.f;
            const uint32_t dst = p_row + (it & 1) * C::P_BYTES;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(s[32 * c + 2 * i]), sl, -mu));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(s[32 * c + 2 * i + 1]), sl, -mu));
                    ls0 += p0; ls1 += p1;
                    pk[i] = pack_bf16(p0, p1);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int cc = c * 4 + q;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((cc ^ rx) << 4)), "r"(pk[4 * q]),
                                 "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                }
            }
            l += ls0 + ls1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(it));
        }
        // epilogue: O / l
        if (n_tiles > 0) {
            mbar_wait(o_done(n_tiles - 1), (uint32_t)((n_tiles - 1) >> 1) & 1u);
            tc_fence_after();
            l += exchange(n_tiles & 1, l);
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        const int i = i0 + row;
        bf16* orow = p.out + ((size_t)b * p.L + (i < p.L ? i : 0)) * p.ldo + h * D;
#pragma unroll 1
        for (int c = half; c < D / 32; c += 2) {
            uint32_t v[32];
            if (n_tiles > 0) {
                tc_ld32(t_lane + C::O_COL + 32 * c, v);
            } else {
#pragma unroll
                for (int k = 0; k < 32; k++) v[k] = 0u;
            }
            if (i < p.L) {
                uint4 ov[4];
                uint32_t* ou = reinterpret_cast<uint32_t*>(ov);
#pragma unroll
                for (int k = 0; k < 16; k++) ou[k] = pack_bf16(__uint_as_float(v[2 * k]) * inv, __uint_as_float(v[2 * k + 1]) * inv);
#pragma unroll
                for (int k = 0; k < 4; k++) reinterpret_cast<uint4*>(orow + 32 * c)[k] = ov[k];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FA_W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int attn_tc_mode() {                     // P3_ATTN_TC=0 forces the mma.sync kernel (A/B testing, cross-check)
    const char* e = getenv("P3_ATTN_TC");
    return (e && e[0] == '0') ? 0 : 1;
}

bool attn_prefill_tc_eligible(const AttnParams& p) {
    if (!attn_tc_mode()) return false;
    if (p.hd != 96 && p.hd != 64) return false;
    if (p.L < 64 || p.past % 128 != 0) return false;
    if ((p.ldq | p.ldk | p.ldv | p.ldo) & 7) return false;
    if (((uintptr_t)p.q | (uintptr_t)p.k | (uintptr_t)p.v | (uintptr_t)p.out | (uintptr_t)p.pool) & 15) return false;
    if ((int64_t)p.B * p.L >= (1ll << 31)) return false;
    return true;
}

template <int D>
static int launch_tc_d(const AttnParams& p, cudaStream_t st) {
    using C = FaCfg<D>;
    CUtensorMap tq, tk, tv, tp;
    const int64_t rows = (int64_t)p.B * p.L;
    CUresult r = tc_encode_2d(&tq, p.q, rows, (int64_t)p.n_heads * D, p.ldq, 128);
    if (r == CUDA_SUCCESS) r = tc_encode_2d(&tk, p.k, rows, (int64_t)p.n_kv * D, p.ldk, 128);
    if (r == CUDA_SUCCESS) r = tc_encode_2d(&tv, p.v, rows, (int64_t)p.n_kv * D, p.ldv, 128);
    if (r == CUDA_SUCCESS) {
        if (p.past > 0) r = tc_encode_2d(&tp, p.pool, (int64_t)1 << 31, D, D, P3_PAGE);   // rows: upper bound, pages are addressed via the block table
        else tp = tk;
    }
    P3_CHECK_ARG(r == CUDA_SUCCESS, "attention_prefill: cuTensorMapEncodeTiled failed (%d)", (int)r);
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(attn_prefill_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        P3_CHECK_ARG(e == cudaSuccess, "attention_prefill: cannot set %d B dynamic smem: %s", C::SMEM, cudaGetErrorString(e));
        set = true;
    }
    dim3 grid((p.L + 127) / 128, p.n_heads, p.B);
    attn_prefill_tc_kernel<D><<<grid, FA_THREADS, C::SMEM, st>>>(tq, tk, tv, tp, p);
    P3_CHECK_LAUNCH("attention_prefill_tc");
    return 0;
}

int launch_prefill_tc(const AttnParams& p, cudaStream_t st) {
    return p.hd == 96 ? launch_tc_d<96>(p, st) : launch_tc_d<64>(p, st);
}
