// Flash attention for many query rows (prefill, ViT) on the 5th-gen tensor cores.
// Same contract as attn_prefill_kernel (attention.cu): softmax(QK^T * scale + mask) V with the
// causal + left-pad predicate of phi.py:550-563, keys [0,past) from the paged pool and keys
// [past,past+L) from the fresh qkv rows (phi.py:454-457; ViT: phi.py:148 with causal = 0).
//
// One CTA = 256 query rows (two 128-row tiles a, b) of one (sequence, head); key tiles of 64 (= one pool page), shared by
// both query tiles (halves the L2 -> SM traffic per flop).
//   warps 0-3 / 4-7 : softmax of query tile a / b — thread r owns query row r (TMEM lane r). The softmax is the critical
//               chain of the kernel (16 K exponentials per 128x128 scores = 1024 MUFU cycles vs 768 cycles of UMMA at
//               head_dim 96), so (1) S is double buffered per query tile: S(j+1) and S(j+2) are computed while tile j is
//               swept and the warpgroup never waits for the tensor pipe in steady state, and (2) S is swept once per tile
//               (32-column tcgen05.ld chunks): P = exp2(S*scale - m*scale) (one FFMA + one MUFU) against the running
//               reference maximum m of the previous tiles, rounded to bf16 and stored back to TMEM over the consumed S
//               columns (P is the A operand of the PV UMMA straight from tensor memory: no shared-memory round trip),
//               while the tile's own maximum is collected in the same sweep. O and l are rescaled lazily
//               (reference grew by more than 2^8: before the next tile; more than 2^64: the tile is redone).
//   warp 8    : TMA producer — Q tiles once, then K and V tiles through two independent 4-stage mbarrier rings.
//               A pool page is one box; the fresh rows come straight from the qkv buffer.
//   warps 9, 10 : MMA issuers of query tile a / b (whole warp converged, elect.sync lane issues; warp 9 also owns the TMEM
//               allocation) — per key tile, once the softmax warpgroup has stored P(j): O += P(j) V(j) (A from TMEM), then
//               S(j+2) = Q K(j+2)^T into the buffer that P(j) occupied (UMMAs execute in issue order).
// TMEM: S_a[2], S_b[2] (64 columns each), O_a, O_b (D columns each). Operands: Q, K as K-major tiles — a 64-dim
// SWIZZLE_128B box plus, for head_dim 96, a 32-dim SWIZZLE_64B box; V consumed in place as an MN-major B
// operand (its [key][dim] layout is already N-contiguous): two SWIZZLE_128B boxes for head_dim 96 so that one N = 96
// UMMA covers the head (the upper half of the second box is never read); P as the TMEM A operand.
#include "attn_common.cuh"
#include "tc_common.cuh"
#include "../../include/phi3_b200.h"
#include <cstdlib>
#include <cstdio>
#include <type_traits>

template <int D>
struct FaCfg {
    static constexpr bool TWO = (D == 96);              // second, 32-dim box
    static constexpr int KT = 64;                       // keys per tile (= one page of the pool)
    static constexpr int QB0 = 128 * 128, QB1 = TWO ? 128 * 64 : 0;     // Q tile: 128 rows, SW128 box (+ SW64 box)
    static constexpr int QTILE = QB0 + QB1;
    static constexpr int KB0 = KT * 128, KB1 = TWO ? KT * 64 : 0;       // K tile: 64 keys
    static constexpr int KTILE = KB0 + KB1;
    static constexpr int VTILE = TWO ? 2 * KB0 : KB0;   // V tile: head_dim 96 takes two SW128 boxes (dims 64..127, upper half unused) so
                                                        // that O += P V is ONE N = 96 UMMA per 16 keys
    static constexpr int STAGES = 4;
    static constexpr int SMEM = 2 * QTILE + STAGES * (KTILE + VTILE) + 1024 + 512;
    // TMEM: S_x[buf] at 128 x + 64 buf (64 fp32 columns; P_x, bf16 pairs, overwrites its first 32 once the row is consumed),
    // O_x at 256 + 128 x
    static constexpr int TMEM_COLS = 512, S_COL = 0, O_COL = 256;
    static constexpr uint32_t IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    static constexpr uint32_t IDESC_QK = IDESC_BASE | ((uint32_t)(KT >> 3) << 17);
    static constexpr uint32_t IDESC_PV = IDESC_BASE | (1u << 16) | ((uint32_t)(D >> 3) << 17);   // B (V) MN-major, N = D
};

struct FaMaps {                                         // [0]: 64-column SW128 box, [1]: 32-column SW64 box (head_dim 96, Q and K)
    CUtensorMap q[2], k[2], v, pool[2];                 // Q boxes have 128 rows, K / V / pool boxes 64; V uses the [0]-type box twice
};

#define FA_THREADS 352
#define FA_W_TMA 8
#define FA_W_MMA 9                                      // warps 9, 10: MMA issuers of query tile a, b
#define FA_RESCALE_LOG2 8.0f
#define FA_PAIR_GROUP 8

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_prefill_tc_kernel(const __grid_constant__ FaMaps tm, AttnParams p, long long* dbg) {
    using C = FaCfg<D>;
    constexpr int ST = C::STAGES, KT = C::KT;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base, sK0 = sQ + 2 * C::QTILE, sV0 = sK0 + ST * C::KTILE;
    const uint32_t bars = sV0 + ST * C::VTILE;
    const uint32_t q_full = bars;
    auto k_full = [&](int s) { return bars + 8u * (1 + s); };
    auto k_empty = [&](int s) { return bars + 8u * (1 + ST + s); };
    auto v_full = [&](int s) { return bars + 8u * (1 + 2 * ST + s); };
    auto v_empty = [&](int s) { return bars + 8u * (1 + 3 * ST + s); };
    auto s_full = [&](int x, int j) { return bars + 8u * (1 + 4 * ST + 2 * x + (j & 1)); };    // S_x(j) complete in buffer j & 1
    // softmax_x(j) done: P_x(j) stored over S_x(j). Alternating pair: with S double buffered the softmax can finish tiles j and
    // j+1 before the MMA warp looks at tile j, and a single barrier would then show an aliased parity (deadlock)
    auto sm_done = [&](int x, int j) { return bars + 8u * (5 + 4 * ST + 2 * x + (j & 1)); };
    auto o_done = [&](int x, int j) { return bars + 8u * (9 + 4 * ST + 2 * x + (j & 1)); };    // O_x += P_x(j) V(j) retired (alternating: rarely waited)
    const uint32_t tmem_slot = bars + 8u * (13 + 4 * ST);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTA order: groups of FA_PAIR_GROUP query-row blocks, heaviest (latest rows) first across ALL (sequence, head)
    // pairs — a per-head heavy-first order leaves the last head's longest CTA for the tail — while the blocks of
    // one group and head stay adjacent so that their K/V reads share L2.
    const int n_pairs = (p.L + 255) / 256, hb = p.n_heads * p.B;
    const int grp = blockIdx.x / (FA_PAIR_GROUP * hb), rem = blockIdx.x - grp * (FA_PAIR_GROUP * hb);
    const int gs = min(FA_PAIR_GROUP, n_pairs - grp * FA_PAIR_GROUP);
    const int bh = rem / gs, pr = grp * FA_PAIR_GROUP + (rem - bh * gs);
    const int i0 = (n_pairs - 1 - pr) * 256, h = bh % p.n_heads, b = bh / p.n_heads;
    const int kvh = h / (p.n_heads / p.n_kv);
    const int past = p.past, s_total = past + p.L;
    const int crow = b / p.row_div;
    const int kv0 = p.kv_start ? p.kv_start[crow] : 0;
    const int n_begin = kv0 / KT;
    const bool act_b = i0 + 128 < p.L;
    int n_x[2];
#pragma unroll
    for (int x = 0; x < 2; x++) {
        int n_end = (s_total + KT - 1) / KT;
        if (p.causal) n_end = min(n_end, (past + min(i0 + 128 * x + 127, p.L - 1)) / KT + 1);
        n_x[x] = max(n_end - n_begin, 0);
    }
    if (!act_b) n_x[1] = 0;
    const int n_max = max(n_x[0], n_x[1]);
#ifdef P3_FA_TIMING   // clock64 stamps of CTA 0's pipeline (slot: 0/1 softmax a/b, 2/3 MMA issue for a/b); printed by the launcher
    const bool dbg_on = dbg && blockIdx.x == 0;
#define FA_T(slot, it, k) do { if (dbg_on && (it) < 64 && (threadIdx.x & 31) == 0) dbg[((slot) * 64 + (it)) * 8 + (k)] = clock64(); } while (0)
#else
#define FA_T(slot, it, k) do { } while (0)
#endif

    FA_T(3, 0, 5);
    if (warp == FA_W_TMA && lane == 0) {
#pragma unroll
        for (int i = 0; i < (C::TWO ? 2 : 1); i++) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.q[i])) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.k[i])) : "memory");
            if (i == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.v)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.pool[i])) : "memory");
        }
    }
    if (warp == FA_W_MMA) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int s = 0; s < ST; s++) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 2); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 2); }   // K/V released by both MMA warps
            for (int x = 0; x < 2; x++) {
                for (int j = 0; j < 2; j++) {
                    mbar_init(s_full(x, j), 1); mbar_init(o_done(x, j), 1);
                    mbar_init(sm_done(x, j), 4);                 // one elected arrival per softmax warp
                }
            }
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
    FA_T(3, 0, 6);
    pdl_trigger();                                               // setup above overlaps the previous kernel's tail
    pdl_wait();                                                  // q/k/v, the pool and `out` belong to earlier kernels until here

    if (warp == FA_W_TMA) {
        // ---------------- TMA producer ----------------
        if (lane == 0 && n_max > 0) {
            mbar_expect_tx(q_full, C::QTILE * (act_b ? 2 : 1));
            for (int x = 0; x < (act_b ? 2 : 1); x++) {
                tma_load_2d(sQ + x * C::QTILE, &tm.q[0], q_full, h * D, b * p.L + i0 + 128 * x);
                if (C::TWO) tma_load_2d(sQ + x * C::QTILE + C::QB0, &tm.q[1], q_full, h * D + 64, b * p.L + i0 + 128 * x);
            }
            const int32_t* bt = p.block_table ? p.block_table + (size_t)crow * p.bt_stride : nullptr;
            auto load_tile = [&](int n, int kv, uint32_t dst, uint32_t bar) {
                // K: SW128 box (dims 0..63) + SW64 box (dims 64..95). V: two SW128 boxes (dims 0..63, 64..127; the columns past
                // the head are the next head's or zero-filled, and no UMMA reads them)
                mbar_expect_tx(bar, kv ? C::VTILE : C::KTILE);
                const bool pool = n * KT < past;                 // one 64-key page of the pool, or 64 fresh rows of the qkv buffer
                const int row = pool ? ((bt[n] * 2 + kv) * p.n_kv + kvh) * P3_PAGE : b * p.L + (n * KT - past);
                const int col = pool ? 0 : kvh * D;
                tma_load_2d(dst, pool ? &tm.pool[0] : (kv ? &tm.v : &tm.k[0]), bar, col, row);
                if (C::TWO) {
                    if (kv) tma_load_2d(dst + C::KB0, pool ? &tm.pool[0] : &tm.v, bar, col + 64, row);
                    else tma_load_2d(dst + C::KB0, pool ? &tm.pool[1] : &tm.k[1], bar, col + 64, row);
                }
            };
            for (int it = 0; it < n_max; it++) {
                const int st = it % ST;
                const uint32_t ph = (uint32_t)(it / ST) & 1u;
                mbar_wait(k_empty(st), ph ^ 1u);
                load_tile(n_begin + it, 0, sK0 + st * C::KTILE, k_full(st));
                mbar_wait(v_empty(st), ph ^ 1u);
                load_tile(n_begin + it, 1, sV0 + st * C::VTILE, v_full(st));
            }
        }
    } else if (warp == FA_W_MMA || warp == FA_W_MMA + 1) {
        // ---------------- MMA issuers: one warp per query tile (whole warp converged; UMMAs / commits by one elected lane).
        // The two chains sweep -> PV -> S are independent; only the K / V stages are shared, and those are released by a
        // commit from BOTH warps for every key tile (a warp whose query tile needs fewer key tiles just commits).
        const int x = warp - FA_W_MMA, nx = n_x[x];
        if (n_max > 0) {
            mbar_wait(q_full, 0);
            // S_x(j) = Q_x K(j)^T into buffer j & 1 (64 fp32 columns)
            auto issue_s = [&](int j) {
                const int st = j % ST;
                const uint32_t d_tmem = tmem_base + C::S_COL + x * 128 + (j & 1) * 64;
                const uint32_t qa = sQ + x * C::QTILE, kb = sK0 + st * C::KTILE;
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
                    tc_mma_bf16(d_tmem, umma_desc_sw128(qa) + 2 * kk, umma_desc_sw128(kb) + 2 * kk, C::IDESC_QK, kk ? 1u : 0u);
                if (C::TWO) {
#pragma unroll
                    for (int kk = 0; kk < 2; kk++)
                        tc_mma_bf16(d_tmem, umma_desc_sw64(qa + C::QB0) + 2 * kk, umma_desc_sw64(kb + C::KB0) + 2 * kk, C::IDESC_QK, 1u);
                }
                tc_commit(s_full(x, j));
            };
            // prologue: the first two score tiles
            for (int j = 0; j < 2 && j < n_max; j++) {
                // (a warp that does not use a stage still waits for it before releasing it: its commits can then never run
                // a ring phase ahead of the other warp's)
                mbar_wait(k_full(j % ST), (uint32_t)(j / ST) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    if (j < nx) issue_s(j);
                    tc_commit(k_empty(j % ST));
                }
                __syncwarp();
            }
            for (int j = 0; j < n_max; j++) {
                const int stv = j % ST;
                mbar_wait(v_full(stv), (uint32_t)(j / ST) & 1u);
                if (j + 2 < n_max) mbar_wait(k_full((j + 2) % ST), (uint32_t)((j + 2) / ST) & 1u);
                if (j < nx) mbar_wait(sm_done(x, j), (uint32_t)(j >> 1) & 1u);   // P_x(j) is in TMEM (exactly one wait per tile, in order)
                tc_fence_after();
                FA_T(2 + x, j, 0);
                if (elect_one()) {
                    if (j < nx) {
                        // O_x += P_x(j) V(j): A from TMEM (lane = row, two bf16 keys per column), V in place as MN-major B
                        const uint32_t d_tmem = tmem_base + C::O_COL + x * 128, a_tmem = tmem_base + C::S_COL + x * 128 + (j & 1) * 64;
                        const uint32_t vb = sV0 + stv * C::VTILE;
#pragma unroll
                        for (int kk = 0; kk < KT / 16; kk++)
                            tc_mma_bf16_ts(d_tmem, a_tmem + 8 * kk, umma_desc_mn_sw128(vb + kk * (16 * 128), C::KB0, 1024), C::IDESC_PV,
                                           (j | kk) ? 1u : 0u);
                        tc_commit(o_done(x, j));
                        // S_x(j+2) reuses the buffer whose P the PV above consumes (UMMAs execute in issue order)
                        if (j + 2 < nx) issue_s(j + 2);
                    }
                    tc_commit(v_empty(stv));
                    if (j + 2 < n_max) tc_commit(k_empty((j + 2) % ST));
                }
                __syncwarp();
                FA_T(2 + x, j, 1);
            }
        }
    } else {
        // ---------------- softmax warpgroup x: thread = query row ----------------
        const int x = warp >> 2, wq = warp & 3;
        const int row = wq * 32 + lane;
        const int nt = n_x[x];
        const int qi = past + i0 + 128 * x + row;
        const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
        const uint32_t t_o = t_lane + C::O_COL + x * 128;
        const float sl = p.scale_log2;
        float m = -INFINITY, l = 0.f;
        float pend = 1.f;                                        // rescale of O_x and l decided after the previous tile
        bool has_pend = false;
        for (int it = 0; it < nt; it++) {
            const int j0 = (n_begin + it) * KT;
            const uint32_t t_s = t_lane + C::S_COL + x * 128 + (it & 1) * 64;
            mbar_wait(s_full(x, it), (uint32_t)(it >> 1) & 1u); // S_x(it) ready (issued two tiles ago: normally no wait)
            tc_fence_after();
            if (wq == 0) FA_T(x, it, 0);
            auto scale_o = [&](float cf) {                       // warp-collective: every lane runs the TMEM round trip
                // O_x must be complete up to tile it-1 and PV_x(it) waits for our P: rare path, so the barrier is only
                // waited here (two alternating barriers keep the parity unambiguous)
                if (it > 0) { mbar_wait(o_done(x, it - 1), (uint32_t)((it - 1) >> 1) & 1u); tc_fence_after(); }
#pragma unroll 1
                for (int c = 0; c < D / 32; c++) {
                    uint32_t v[32];
                    tc_ld32(t_o + 32 * c, v);
#pragma unroll
                    for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * cf);
                    tc_st32(t_o + 32 * c, v);
                }
            };
            if (__any_sync(0xffffffffu, has_pend)) scale_o(has_pend ? pend : 1.f);
            if (has_pend) { l *= pend; has_pend = false; }
            const bool need_mask = (j0 < kv0) || (j0 + KT - 1 >= s_total) || (p.causal && j0 + KT - 1 > past + i0 + 128 * x + wq * 32);
            uint32_t pk[32];                                     // the P row, bf16 pairs
            // One sweep over the S row in TMEM (two 32-column chunks, both in flight): row maximum and / or
            // P = exp2(S * scale - mu) rounded to bf16.
            auto sweep = [&](auto do_max, auto do_exp, float mu, float& mx_out, float& ls_out) {
                constexpr bool DM = decltype(do_max)::value, DE = decltype(do_exp)::value;
                uint32_t buf[2][32];
                tc_ld32_nowait(t_s, buf[0]);
                tc_ld32_nowait(t_s + 32, buf[1]);
                tc_wait_ld();
                tc_reg_fence32(buf[0]);
                tc_reg_fence32(buf[1]);
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY, ls0 = 0.f, ls1 = 0.f;
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    uint32_t* v = buf[c];
                    if (need_mask) {
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const int j = j0 + 32 * c + i;
                            const bool ok = (j >= kv0) && (j < s_total) && (!p.causal || j <= qi);
                            if (!ok) v[i] = 0xff800000u;
                        }
                    }
                    if (DM) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            mx0 = fmaxf(mx0, __uint_as_float(v[i])); mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
                            mx2 = fmaxf(mx2, __uint_as_float(v[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
                        }
                    }
                    if (DE) {
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), sl, -mu));
                            const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), sl, -mu));
                            ls0 += p0; ls1 += p1;
                            pk[16 * c + i] = pack_bf16(p0, p1);
                        }
                    }
                }
                if (DM) mx_out = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                if (DE) ls_out = ls0 + ls1;
            };
            // Streaming softmax: P is computed against the reference maximum m carried over from the previous
            // tiles while this tile's own maximum is found in the same sweep (TMEM is read once per tile). If the
            // tile's maximum turns out to exceed the reference by more than 2^64 (or there was no reference yet),
            // the tile is redone against its own maximum; smaller growth (> 2^8) only schedules a rescale of O
            // and l for the next tile. O / l is invariant under the reference, so the result is exact either way.
            float mx = -INFINITY, ls = 0.f;
            bool redo = true;
            if (it > 0) {
                sweep(std::true_type{}, std::true_type{}, (m == -INFINITY) ? 0.f : m * sl, mx, ls);
                redo = (m == -INFINITY) ? (mx != -INFINITY) : ((mx - m) * sl > 64.f);
            } else {
                sweep(std::true_type{}, std::false_type{}, 0.f, mx, ls);
            }
            if (__any_sync(0xffffffffu, redo)) {
                float corr = 1.f;
                if (redo) { corr = (m == -INFINITY) ? 0.f : ex2_approx((m - mx) * sl); m = mx; }
                if (it > 0) scale_o(corr);
                l *= corr;
                sweep(std::false_type{}, std::true_type{}, (m == -INFINITY) ? 0.f : m * sl, mx, ls);
            }
            l += ls;
            if (wq == 0) FA_T(x, it, 2);
            const float grow = (mx - m) * sl;                    // NaN (no keys yet) compares false
            if (grow > FA_RESCALE_LOG2) { pend = ex2_approx(-grow); has_pend = true; m = mx; }
            // P_x(it) over the first 32 columns of its own S buffer (the row is fully consumed; S_x(it+2) is issued after PV_x(it))
            tc_st32(t_s, pk);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sm_done(x, it));
            if (wq == 0) FA_T(x, it, 3);
        }
        // ---- epilogue: O / l
        if (x == 0 || act_b) {
            if (nt > 0) {
                mbar_wait(o_done(x, nt - 1), (uint32_t)((nt - 1) >> 1) & 1u);   // the last PV_x has retired
                tc_fence_after();
            }
            const float inv = l > 0.f ? 1.f / l : 0.f;
            const int i = i0 + 128 * x + row;
            bf16* orow = p.out + ((size_t)b * p.L + (i < p.L ? i : 0)) * p.ldo + h * D;
#pragma unroll 1
            for (int c = 0; c < D / 32; c++) {
                uint32_t v[32];
                if (nt > 0) {
                    tc_ld32(t_o + 32 * c, v);
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
    }
    if (warp < 4) FA_T(3, 1 + warp, 5);
    tc_fence_before();
    __syncthreads();
    FA_T(3, 0, 7);
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
    if (p.L < 64 || p.past % 64 != 0) return false;
    if ((p.ldq | p.ldk | p.ldv | p.ldo) & 7) return false;
    if (((uintptr_t)p.q | (uintptr_t)p.k | (uintptr_t)p.v | (uintptr_t)p.out | (uintptr_t)p.pool) & 15) return false;
    if ((int64_t)p.B * p.L >= (1ll << 31)) return false;
    return true;
}

template <int D>
static int launch_tc_d(const AttnParams& p, cudaStream_t st) {
    using C = FaCfg<D>;
    FaMaps tm;
    const int64_t rows = (int64_t)p.B * p.L;
    CUresult r = CUDA_SUCCESS;
    for (int i = 0; i < (C::TWO ? 2 : 1) && r == CUDA_SUCCESS; i++) {
        const int bc = i ? 32 : 64;
        r = tc_encode_2d(&tm.q[i], p.q, rows, (int64_t)p.n_heads * D, p.ldq, 128, bc);
        if (r == CUDA_SUCCESS) r = tc_encode_2d(&tm.k[i], p.k, rows, (int64_t)p.n_kv * D, p.ldk, C::KT, bc);
        if (r == CUDA_SUCCESS && i == 0) r = tc_encode_2d(&tm.v, p.v, rows, (int64_t)p.n_kv * D, p.ldv, C::KT, 64);
        if (r == CUDA_SUCCESS) {
            // pool rows: an upper bound — pages are addressed through the block table
            if (p.past > 0) r = tc_encode_2d(&tm.pool[i], p.pool, (int64_t)1 << 31, D, D, P3_PAGE, bc);
            else tm.pool[i] = tm.k[i];
        }
    }
    if (!C::TWO) { tm.q[1] = tm.q[0]; tm.k[1] = tm.k[0]; tm.pool[1] = tm.pool[0]; }
    P3_CHECK_ARG(r == CUDA_SUCCESS, "attention_prefill: cuTensorMapEncodeTiled failed (%d)", (int)r);
    static P3DevFlags flags; bool& set = flags.cur();
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(attn_prefill_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        P3_CHECK_ARG(e == cudaSuccess, "attention_prefill: cannot set %d B dynamic smem: %s", C::SMEM, cudaGetErrorString(e));
        set = true;
    }
    dim3 grid(((p.L + 255) / 256) * p.n_heads * p.B);
    long long* dbg = nullptr;
#ifdef P3_FA_TIMING
    const bool want_dbg = getenv("P3_FA_DBG") != nullptr;
    if (want_dbg) { cudaMalloc(&dbg, 4 * 64 * 8 * 8); cudaMemset(dbg, 0, 4 * 64 * 8 * 8); }
#endif
    p3_launch_pdl(attn_prefill_tc_kernel<D>, grid, dim3(FA_THREADS), (size_t)C::SMEM, st, tm, p, dbg);
    P3_CHECK_LAUNCH("attention_prefill_tc");
#ifdef P3_FA_TIMING
    if (want_dbg) {
        static long long h[4 * 64 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        long long t0 = h[0];
        auto H = [&](int s, int it, int k) { return h[(s * 64 + it) * 8 + k] - t0; };
        for (int it = 0; it < 64 && h[(0 * 64 + it) * 8]; it++) {
            printf("it %2d | smA: sfull %7lld swept %7lld arrived %7lld | smB: %7lld %7lld %7lld | mma top %7lld kfull %7lld | A: woke %7lld vfull %7lld pv %7lld s %7lld commit %7lld | B: %7lld %7lld %7lld %7lld %7lld\n", it,
                   H(0, it, 0), H(0, it, 2), H(0, it, 3), H(1, it, 0), H(1, it, 2), H(1, it, 3), H(2, it, 5), H(2, it, 6),
                   H(2, it, 0), H(2, it, 2), H(2, it, 3), H(2, it, 4), H(2, it, 1), H(3, it, 0), H(3, it, 2), H(3, it, 3), H(3, it, 4), H(3, it, 1));
        }
        printf("entry %lld setup %lld | softmax A warps done %lld %lld %lld %lld | end %lld\n", H(3, 0, 5), H(3, 0, 6), H(3, 1, 5), H(3, 2, 5), H(3, 3, 5), H(3, 4, 5), H(3, 0, 7));
        fflush(stdout);
    }
#endif
    return 0;
}

int launch_prefill_tc(const AttnParams& p, cudaStream_t st) {
    return p.hd == 96 ? launch_tc_d<96>(p, st) : launch_tc_d<64>(p, st);
}
