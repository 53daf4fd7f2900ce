// Persistent decode-layer kernel: the weight stream of a whole decoder layer in ONE launch.
//
// Replaces, for M <= 8 decode rows, the chain of per-matrix skinny GEMMs
//     o_proj (+residual) -> RMSNorm -> gate_up_proj (+SwiGLU) -> down_proj (+residual) -> RMSNorm -> qkv_proj
//     (+SuRoPE, +paged KV write)   |   ... -> RMSNorm -> lm_head
// (phi.py:437-438, 442-453, 460, 465-471, 478-485, 604-608) by one launch of <#SM> persistent CTAs that walk the
// phases with grid-wide barriers in between. The op is a pure HBM weight stream (7.4 GB per decode step), so the
// design goal is that HBM never idles at an op boundary:
//
//   * weights are re-packed once at load time (p3_mega_pack) into "stream order": for every 16-row tile, K-block and
//     consumer warp one contiguous segment whose 512-byte pieces are exactly the per-lane A fragments of
//     mma.sync.m16n8k16 (rows of W are the MMA "M", the <= 8 tokens are "N"), so a warp consumes its stream with
//     conflict-free 16-byte ld.shared and no transposes;
//   * every consumer warp owns a private ring of MG_RSLOTS x 4 KB shared-memory slots filled by cp.async.bulk
//     (mbarrier complete_tx, L2 evict-first) that it issues itself one ring ahead of its read position. The ring
//     position runs through ALL phases of the launch, i.e. while a warp waits at a grid barrier (or, under
//     programmatic dependent launch, for the previous kernel to finish) 192 KB per SM = 28 MB per GPU of the NEXT
//     phase's weights are already in flight — more than HBM delivers during the barrier;
//   * tiles are mapped to CTAs by a host-built schedule that balances CUMULATIVE bytes per CTA at every phase
//     boundary (so all CTAs reach each barrier together) instead of per-matrix tile counts;
//   * activations (<= 8 x 8192 bf16) live in L2: x fragments are loaded once per phase into registers
//     (ld.global.cg, after the barrier), RMSNorm is applied to them in registers from per-tile sum-of-squares
//     partials written by the producing phase (fixed summation order: deterministic).
//
// Roofline: HBM. Algorithmic bytes per launch = sum over phases of N*K*2 (226.5 MB per Phi-3.5 layer).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/phi3_b200.h"

#define MG_WARPS 8
#define MG_THREADS (MG_WARPS * 32 + 32)      // 8 consumer warps + 1 producer warp
#define MG_CONSUMERS (MG_WARPS * 32)
#ifndef MG_SLOT
#define MG_SLOT 4096
#endif
#ifndef MG_RSLOTS
#define MG_RSLOTS 6
#endif
#define MG_XF 32                       // max 16-wide k blocks per warp and K-block  (K-block <= 8 * 32 * 16 = 4096)
#define MG_KBLOCK 4096
#define MG_MAX_PART 8                  // tiles per CTA of a multi-K-block phase (partial sums parked in smem)
#define MG_RED_STRIDE 136              // floats per (warp, m-tile) in the reduction scratch: n * 17 + r

#define MG_RING_BYTES (MG_WARPS * MG_RSLOTS * MG_SLOT)
#define MG_RED_BYTES (2 * MG_WARPS * 2 * MG_RED_STRIDE * 4)
#define MG_PART_BYTES (MG_MAX_PART * 2 * 128 * 4)
#define MG_SMEM (MG_RING_BYTES + MG_RED_BYTES + MG_PART_BYTES + 2 * MG_WARPS * MG_RSLOTS * 8)

__host__ __device__ __forceinline__ int mg_mt(int kind) { return kind == P3_MEGA_RESID ? 1 : 2; }

// k offset (inside a warp's K slice) of the 4 consecutive elements lane t contributes to k block j. Blocks are paired so that
// one 16-byte load of X serves two MMAs: pair p = j/2 covers 32 k, lane t owns [32p + 8t, +8), first half -> block 2p,
// second half -> block 2p+1; an unpaired last block (odd nkb_w) owns [16j + 4t, +4).
__host__ __device__ __forceinline__ int mg_koff(int j, int t, int nkb_w) {
    return (j | 1) < nkb_w ? 32 * (j >> 1) + 8 * t + 4 * (j & 1) : 16 * j + 4 * t;
}

// first W row of m-tile `mt` of tile `ti` (the same map drives p3_mega_pack and the epilogues)
__host__ __device__ __forceinline__ int mg_tile_row(int kind, int ti, int mt, int N, int n_heads, int n_kv, int hd) {
    if (kind == P3_MEGA_RESID) return 16 * ti;
    if (kind == P3_MEGA_SWIGLU) return mt * (N / 2) + 16 * ti;                  // gate rows | up rows (phi.py:470)
    if (kind == P3_MEGA_QKV_ROPE) {
        const int gpr = hd / 32, n_rope = (n_heads + n_kv) * gpr;
        if (ti < n_rope) return (ti / gpr) * hd + 16 * (ti % gpr) + mt * (hd / 2);   // a rotary pair (d, d + hd/2) meets in one tile
        return (n_heads + n_kv) * hd + 32 * (ti - n_rope) + 16 * mt;
    }
    return 32 * ti + 16 * mt;                                                   // P3_MEGA_F32
}

// ------------------------------------------------------------------------------------------------------------------
// p3_mega_pack: W [N, K] bf16 (nn.Linear layout) -> stream order [kb][tile][warp][j][mt][lane][8 bf16]
// lane (g, t) of k block j holds {W[g][k,k+1], W[g+8][k,k+1], W[g][k+2,k+3], W[g+8][k+2,k+3]}, k = mg_koff(j, t), = a0..a3 of
// the m16n8k16 A fragment under a k permutation (logical 2t+e -> k+e, logical 2t+8+e -> k+2+e) that lets the X operand of
// two k blocks be fetched with one 16-byte load per lane.
// ------------------------------------------------------------------------------------------------------------------
__global__ void mega_pack_kernel(const bf16* __restrict__ W, uint4* __restrict__ out, int kind, int N, int K, int n_heads,
                                 int n_kv, int hd, int64_t total) {
    const int MT = mg_mt(kind), T = N / (16 * MT);
    const int n_kblk = (K + MG_KBLOCK - 1) / MG_KBLOCK, nkb_w = K / (n_kblk * 128);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int lane = r % 32; r /= 32;
        const int mt = r % MT; r /= MT;
        const int j = r % nkb_w; r /= nkb_w;
        const int warp = r % MG_WARPS; r /= MG_WARPS;
        const int ti = r % T; r /= T;
        const int kb = (int)r;
        const int g = lane >> 2, t = lane & 3;
        const int row0 = mg_tile_row(kind, ti, mt, N, n_heads, n_kv, hd) + g;
        const int k0 = (kb * MG_WARPS + warp) * nkb_w * 16 + mg_koff(j, t, nkb_w);
        const uint2 lo = *reinterpret_cast<const uint2*>(W + (size_t)row0 * K + k0);          // W[g][4t..4t+3]
        const uint2 hi = *reinterpret_cast<const uint2*>(W + (size_t)(row0 + 8) * K + k0);    // W[g+8][4t..4t+3]
        out[i] = make_uint4(lo.x, hi.x, lo.y, hi.y);
    }
}

static int mega_dims_ok(int kind, int N, int K, int n_heads, int n_kv, int hd) {
    const int MT = mg_mt(kind);
    if (N <= 0 || K <= 0 || N % (16 * MT) != 0) return 0;
    const int n_kblk = (K + MG_KBLOCK - 1) / MG_KBLOCK;
    if (K % (n_kblk * 128) != 0 || K / (n_kblk * 128) > MG_XF) return 0;
    if (kind == P3_MEGA_QKV_ROPE && (hd % 32 != 0 || N != (n_heads + 2 * n_kv) * hd)) return 0;
    return 1;
}

extern "C" int p3_mega_pack(const void* W, void* out, int kind, int N, int K, int n_heads, int n_kv, int hd, cudaStream_t st) {
    P3_CHECK_ARG(kind >= P3_MEGA_RESID && kind <= P3_MEGA_F32, "mega_pack: unknown phase kind %d", kind);
    P3_CHECK_ARG(mega_dims_ok(kind, N, K, n_heads, n_kv, hd), "mega_pack: unsupported shape N=%d K=%d kind=%d", N, K, kind);
    const int64_t total = (int64_t)N * K / 8;
    mega_pack_kernel<<<(unsigned)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192), 256, 0, st>>>(
        (const bf16*)W, (uint4*)out, kind, N, K, n_heads, n_kv, hd, total);
    P3_CHECK_LAUNCH("mega_pack");
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// the persistent kernel
// ------------------------------------------------------------------------------------------------------------------
static_assert(sizeof(p3_mega_phase) == 104 && sizeof(p3_mega_args) == 520, "p3_mega_args layout is mirrored by ctypes in mega.py");
#include <type_traits>
#include <utility>
template <int N, typename Fn, int... Is>
__device__ __forceinline__ void mg_for_slots_impl(Fn&& fn, std::integer_sequence<int, Is...>) {
    (fn(std::integral_constant<int, Is>{}), ...);
}
template <int N, typename Fn>
__device__ __forceinline__ void mg_for_slots(Fn&& fn) { mg_for_slots_impl<N>(fn, std::make_integer_sequence<int, N>{}); }

struct MgDerived { int MT, T, n_kblk, nkb_w; uint32_t seg; };

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ float mg_silu(float x) { return x / (1.f + __expf(-x)); }

#define MG_TILE_CACHE 32               // per-phase tile ids of this CTA kept in shared memory (more: read from global)
#ifndef MG_PF_ITEMS
#define MG_PF_ITEMS 0                  // L2 prefetch runs this many items (96-192 KB each) ahead of the ring
#endif

__device__ __forceinline__ void mg_arrive(unsigned* ctr) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}

// Rule for everything below: the memory system is kept saturated with ~50 MB of outstanding weight requests, so a demand
// load that MISSES L2 on a consumer's critical path waits for that whole queue (several microseconds). Consumers therefore
// only ever load (a) activations written moments ago (L2 hits), (b) data that arrives through their ring in stream order
// (weights and the RMSNorm gains), (c) values fetched at kernel start, before the flood (rope table row, page id).
__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(const __grid_constant__ p3_mega_args P) {
    extern __shared__ __align__(1024) uint8_t mg_smem[];
    __shared__ MgDerived s_d[P3_MEGA_MAX_PHASES];
    __shared__ int s_first[P3_MEGA_MAX_PHASES], s_cnt[P3_MEGA_MAX_PHASES];
    __shared__ int s_tiles[P3_MEGA_MAX_PHASES][MG_TILE_CACHE];
    __shared__ const uint8_t* s_wp[P3_MEGA_MAX_PHASES];
    __shared__ float s_ss[MG_WARPS][8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int cta = blockIdx.x;
    const int M = P.M;

    uint8_t* ring_g = mg_smem + (size_t)(warp & 7) * (MG_RSLOTS * MG_SLOT);
    float* red = reinterpret_cast<float*>(mg_smem + MG_RING_BYTES);
    float* part = reinterpret_cast<float*>(mg_smem + MG_RING_BYTES + MG_RED_BYTES);
    // mbarriers: full[w][slot] (bulk copy landed) and empty[w][slot] (consumer warp w has the slot's fragments in registers)
    const uint32_t bars_all = smem_u32(mg_smem + MG_RING_BYTES + MG_RED_BYTES + MG_PART_BYTES);
    const uint32_t bars = bars_all + (warp & 7) * (MG_RSLOTS * 8);
    const uint32_t ebars = bars + MG_WARPS * MG_RSLOTS * 8;

    if (warp < P.n_phases) {                                    // warp w caches the schedule of phase w
        const p3_mega_phase& ph = P.ph[warp];
        const int first = ph.cta_off[cta], cnt = ph.cta_off[cta + 1] - first;
        if (lane == 0) {
            MgDerived d;
            d.MT = mg_mt(ph.kind); d.T = ph.N / (16 * d.MT);
            d.n_kblk = (ph.K + MG_KBLOCK - 1) / MG_KBLOCK; d.nkb_w = ph.K / (d.n_kblk * 128);
            d.seg = (uint32_t)d.nkb_w * d.MT * 512u;
            s_d[warp] = d;
            s_first[warp] = first;
            s_cnt[warp] = cnt;
            s_wp[warp] = reinterpret_cast<const uint8_t*>(ph.wp);
        }
        if (lane < cnt && lane < MG_TILE_CACHE) s_tiles[warp][lane] = ph.tile_ids[first + lane];
    }
    if (warp < MG_WARPS && lane == 0) {
#pragma unroll
        for (int s = 0; s < MG_RSLOTS; s++) { mbar_init(bars + s * 8, 1); mbar_init(ebars + s * 8, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_trigger();
    auto tile_of = [&](int p, int li) -> int {
        return li < MG_TILE_CACHE ? s_tiles[p][li] : P.ph[p].tile_ids[s_first[p] + li];
    };

    if (warp == MG_WARPS) {
        // ================= producer warp: free-running weight stream (weights are immutable: no dependency wait) ===========
        // Walks (phase, K-block, tile) exactly like the consumers. The 8 warps' segments of an item are adjacent in memory
        // (seg bytes apart), so lane w < 8 feeds ring w with 4 KB pieces. A normed phase starts with one extra piece per ring:
        // that warp's slice of the RMSNorm gains. Whole items are pulled into L2 MG_PF_ITEMS items ahead of the ring, so HBM
        // keeps streaming (into L2) while the consumers sit in a grid barrier with full rings.
        const uint64_t pol = l2_evict_first_policy();
        const uint32_t my_ring = smem_u32(mg_smem) + (lane & 7) * (MG_RSLOTS * MG_SLOT);
        const uint32_t my_full = bars_all + (lane & 7) * (MG_RSLOTS * 8), my_empty = my_full + MG_WARPS * MG_RSLOTS * 8;
        uint32_t slot = 0, epar = 1;                            // a fresh empty barrier passes a parity-1 wait
        auto put = [&](const uint8_t* src, uint32_t bytes) {    // lanes 0-7: one piece into ring `lane`
            if (lane < MG_WARPS) {
                mbar_wait(my_empty + slot * 8, epar);
                mbar_expect_tx(my_full + slot * 8, bytes);
                bulk_g2s(my_ring + slot * MG_SLOT, src, bytes, my_full + slot * 8, pol);
            }
            __syncwarp();
            if (++slot == MG_RSLOTS) { slot = 0; epar ^= 1; }
        };
        // item-granular look-ahead cursor for the L2 prefetch
        int lp = 0, lkb = 0, lli = -1;
        auto l2_next_item = [&]() {
            for (;;) {
                if (lp >= P.n_phases) return;
                if (++lli < s_cnt[lp]) break;
                lli = -1;
                if (++lkb >= s_d[lp].n_kblk) { lkb = 0; lp++; }
            }
            const MgDerived d = s_d[lp];
            const uint32_t item = MG_WARPS * d.seg, per = (item / 32 + 15) & ~15u;      // bytes per lane, 16-byte granular
            const uint8_t* base = s_wp[lp] + (size_t)(lkb * d.T + tile_of(lp, lli)) * item;
            const uint32_t off = lane * per;
            if (off < item)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(min(per, item - off)) : "memory");
        };
        bool first_item = true;
        for (int p = 0; p < P.n_phases; p++) {
            const MgDerived d = s_d[p];
            const int n_mine = s_cnt[p];
            if (n_mine == 0) continue;
            const uint8_t* nw = reinterpret_cast<const uint8_t*>(P.ph[p].norm_w);
            const uint32_t kpw_b = d.nkb_w * 32;                // bytes of a warp's K slice of bf16 gains
            for (int kb = 0; kb < d.n_kblk; kb++) {
                if (nw) put(nw + (size_t)(kb * MG_WARPS + (lane & 7)) * kpw_b, kpw_b);
                for (int li = 0; li < n_mine; li++) {
                    const uint8_t* base = s_wp[p] + ((size_t)(kb * d.T + tile_of(p, li)) * MG_WARPS + (lane & 7)) * d.seg;
                    if (first_item) {                           // ring first, then the look-ahead (skipping what the ring holds)
                        first_item = false;
                        for (uint32_t off = 0; off < d.seg; off += MG_SLOT) put(base + off, min((uint32_t)MG_SLOT, d.seg - off));
                        lp = p; lkb = kb; lli = li;             // look-ahead starts behind this item
#pragma unroll 1
                        for (int i = 0; i < MG_PF_ITEMS; i++) l2_next_item();
                        continue;
                    }
                    l2_next_item();
                    for (uint32_t off = 0; off < d.seg; off += MG_SLOT) put(base + off, min((uint32_t)MG_SLOT, d.seg - off));
                }
            }
        }
        return;
    }

    // ================= consumer warps =================
    auto cta_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(MG_CONSUMERS) : "memory"); };   // consumers only
    uint32_t cslot = 0, cpar = 0;                               // ring position: slot and its phase parity
    auto release_slot = [&]() {                                 // every lane has the slot's data in registers: hand it back
        __syncwarp();
        if (lane == 0) mbar_arrive(ebars + cslot * 8);
        if (++cslot == MG_RSLOTS) { cslot = 0; cpar ^= 1; }
    };

    // ---- QKV phase (always the last one): this thread's rope factors and page id are requested now and used ~50 us later
    // (they may miss L2; nothing waits on them until the epilogue of the last phase)
    const int last_kind = P.ph[P.n_phases - 1].kind;
    float rope_cs[3] = {1.f, 1.f, 1.f}, rope_sn[3] = {0.f, 0.f, 0.f};
    int kv_page = 0;
    pdl_wait();
    const int past = P.past_dev ? __ldcg(P.past_dev) : P.past;
    if (last_kind == P3_MEGA_QKV_ROPE && tid < 128 && (tid >> 4) < M) {
        const int n = tid >> 4, r = tid & 15, half = P.hd / 2;
        const size_t tix = (size_t)n * P.tab_bstride + (size_t)past * half + r;
#pragma unroll
        for (int q = 0; q < 3; q++)
            if (q * 16 + r < half) { rope_cs[q] = __ldg(P.cosT + tix + q * 16); rope_sn[q] = __ldg(P.sinT + tix + q * 16); }
        kv_page = __ldg(P.block_table + (size_t)n * P.bt_stride + past / P3_PAGE);
    }

    unsigned n_bar = 0;
    int red_buf = 0;
    // optional instrumentation (P.dbg, tools/mega_trace.py): per CTA and phase, SM-clock stamps
    //   [0] phase start (barrier arrive)  [1] barrier released  [2] warp 0 has its X fragments  [3] last item done
    //   [4] cycles warp 0 spent waiting for weight data   [5] same, warp 7
    long long* dbg = P.dbg ? P.dbg + (size_t)cta * (P3_MEGA_MAX_PHASES * 8) : nullptr;
    long long wait_cyc = 0, sync_cyc = 0, epi_cyc = 0;
    for (int p = 0; p < P.n_phases; p++) {
        if (dbg && tid == 0) dbg[p * 8 + 0] = clock64();
        if (p > 0) {                                            // grid barrier: everything phase p reads has been written
            cta_sync();
            if (tid == 0) {
                mg_arrive(P.sync);                              // release: orders the CTA's writes (cumulative through bar.sync)
                const unsigned target = ++n_bar * gridDim.x;
#ifndef MG_NO_BARRIER_WAIT                                               // (timing experiment only: results are garbage without the wait)
                unsigned v, spins = 0;
                do {
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.sync) : "memory");
                    if (v < target && ++spins > (1u << 26)) __trap();      // a CTA never arrived: fail loudly instead of hanging the GPU
                } while (v < target);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
#else
                (void)target;
#endif
            }
            cta_sync();
        }
        const p3_mega_phase& ph = P.ph[p];
        const MgDerived d = s_d[p];
        const int n_mine = s_cnt[p];
        if (dbg && tid == 0) dbg[p * 8 + 1] = clock64();
        wait_cyc = sync_cyc = epi_cyc = 0;
        if (n_mine == 0) continue;                              // (uniform per CTA)
        const bf16* X = reinterpret_cast<const bf16*>(ph.x);
        const bool normed = ph.norm_w != nullptr;
        const int n_pair = d.nkb_w >> 1;

        for (int kb = 0; kb < d.n_kblk; kb++) {
            // ---- this warp's X fragments for the K-block, in registers: one 16-byte L2 load per pair of k blocks
            uint32_t xf[MG_XF][2];
            const int kbase = (kb * MG_WARPS + warp) * d.nkb_w * 16;
            const bf16* xrow = X + (size_t)g * ph.ldx + kbase;
#pragma unroll
            for (int pr = 0; pr < MG_XF / 2; pr++) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (pr < n_pair && g < M) v = __ldcg(reinterpret_cast<const uint4*>(xrow + 32 * pr + 8 * t));
                xf[2 * pr][0] = v.x; xf[2 * pr][1] = v.y; xf[2 * pr + 1][0] = v.z; xf[2 * pr + 1][1] = v.w;
            }
            if (d.nkb_w & 1) {                                  // unpaired last k block (small K only)
                const int j = d.nkb_w - 1;
                uint2 v = make_uint2(0u, 0u);
                if (g < M) v = __ldcg(reinterpret_cast<const uint2*>(xrow + 16 * j + 4 * t));
#pragma unroll
                for (int jj = 0; jj < MG_XF; jj += 2)
                    if (jj == j) { xf[jj][0] = v.x; xf[jj][1] = v.y; }
            }
            if (normed) {
                // ---- RMSNorm (phi.py:478-479): rs of row g from the producer's per-tile partial sums, summed by the CTA in a
                // fixed order (deterministic): thread -> partial c = tid/2 (+128 i), rows 4q..4q+3. The X loads above are in flight.
                if (kb == 0) {
                    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int q = tid & 1;
                    if (ph.ss_in) {
                        for (int c = tid >> 1; c < ph.n_ss_in; c += MG_CONSUMERS / 2) {
                            const float4 v = __ldcg(reinterpret_cast<const float4*>(ph.ss_in + (size_t)c * 16 + 4 * q));
                            a4.x += v.x; a4.y += v.y; a4.z += v.z; a4.w += v.w;
                        }
                    } else {                                    // no partials: thread -> 8-element chunks of x
                        for (int c = tid >> 1; c < ph.K / 8; c += MG_CONSUMERS / 2) {
                            float* av = reinterpret_cast<float*>(&a4);
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                if (4 * q + i < M) {
                                    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(X + (size_t)(4 * q + i) * ph.ldx) + c);
                                    const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
                                    for (int k = 0; k < 4; k++) { const float2 f = unpack_bf16(u[k]); av[i] += f.x * f.x + f.y * f.y; }
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int o = 2; o < 32; o <<= 1) {
                        a4.x += __shfl_xor_sync(0xffffffffu, a4.x, o); a4.y += __shfl_xor_sync(0xffffffffu, a4.y, o);
                        a4.z += __shfl_xor_sync(0xffffffffu, a4.z, o); a4.w += __shfl_xor_sync(0xffffffffu, a4.w, o);
                    }
                    if (lane < 2) *reinterpret_cast<float4*>(&s_ss[warp][4 * lane]) = a4;
                    cta_sync();
                }
                float sum = 0.f;
#pragma unroll
                for (int w = 0; w < MG_WARPS; w++) sum += s_ss[w][g];
                const float rs = rsqrtf(sum / (float)ph.K + P.eps);
                // the gains of this warp's K slice arrive through the ring, ahead of the K-block's weights
                mbar_wait(bars + cslot * 8, cpar);
                const bf16* nws = reinterpret_cast<const bf16*>(ring_g + cslot * MG_SLOT);
#pragma unroll
                for (int pr = 0; pr < MG_XF / 2; pr++) {
                    if (pr < n_pair) {
                        const uint4 w = *reinterpret_cast<const uint4*>(nws + 32 * pr + 8 * t);
                        const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            uint32_t& xr = xf[2 * pr + (e >> 1)][e & 1];
                            const float2 a = unpack_bf16(xr), ww = unpack_bf16(wv[e]);
                            xr = pack_bf16(a.x * rs * ww.x, a.y * rs * ww.y);
                        }
                    }
                }
                if (d.nkb_w & 1) {
                    const int j = d.nkb_w - 1;
                    const uint2 w = *reinterpret_cast<const uint2*>(nws + 16 * j + 4 * t);
#pragma unroll
                    for (int jj = 0; jj < MG_XF; jj += 2)
                        if (jj == j) {
                            const float2 a = unpack_bf16(xf[jj][0]), b = unpack_bf16(xf[jj][1]);
                            const float2 wa = unpack_bf16(w.x), wb = unpack_bf16(w.y);
                            xf[jj][0] = pack_bf16(a.x * rs * wa.x, a.y * rs * wa.y);
                            xf[jj][1] = pack_bf16(b.x * rs * wb.x, b.y * rs * wb.y);
                        }
                }
                release_slot();
            }
            if (dbg && tid == 0 && kb == 0) dbg[p * 8 + 2] = clock64();
            const bool last_kb = (kb == d.n_kblk - 1);
            const int F = d.nkb_w * d.MT;                       // 512-byte fragments per item

            for (int li = 0; li < n_mine; li++) {
                const int ti = tile_of(p, li);
                // RESID epilogue: the residual value this thread will add is known up front -> fetch it under the MMAs
                float resid_pref = 0.f;
                if (last_kb && ph.kind == P3_MEGA_RESID && tid < 128 && (tid >> 4) < M) {
                    const unsigned short raw = __ldcg(reinterpret_cast<const unsigned short*>(ph.out) +
                                                      (size_t)(tid >> 4) * ph.ldo + 16 * ti + (tid & 15));
                    resid_pref = __bfloat162float(__ushort_as_bfloat16(raw));
                }
                float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                // ---- stream the item slot by slot: wait, 8 x ld.shared.v4, 8 x MMA, hand the slot back.
                // MT == 2: fragment f = 2j + mt feeds acc[mt]. MT == 1: fragment f = j feeds acc[j & 1] (two MMA chains).
                auto slot_group = [&](auto S_, auto MT2_) {
                    constexpr int S = decltype(S_)::value;
                    constexpr bool MT2 = decltype(MT2_)::value;
                    if (8 * S >= F) return;
                    if (dbg) { const long long t0 = clock64(); mbar_wait(bars + cslot * 8, cpar); wait_cyc += clock64() - t0; }
                    else mbar_wait(bars + cslot * 8, cpar);
                    const uint8_t* sb = ring_g + cslot * MG_SLOT + lane * 16;
                    if (8 * S + 8 <= F) {                       // full slot
                        uint4 a[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const uint4*>(sb + i * 512);
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int f = 8 * S + i, j = MT2 ? f / 2 : f;
                            const uint32_t av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
                            mma_bf16_16816(acc[f & 1], av, xf[j][0], xf[j][1]);
                        }
                    } else {                                    // partial last slot of the item (small K only)
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int f = 8 * S + i, j = MT2 ? f / 2 : f;
                            if (f < F) {
                                const uint4 a = *reinterpret_cast<const uint4*>(sb + i * 512);
                                const uint32_t av[4] = {a.x, a.y, a.z, a.w};
                                mma_bf16_16816(acc[f & 1], av, xf[j][0], xf[j][1]);
                            }
                        }
                    }
                    release_slot();
                };
                if (d.MT == 1) {
                    mg_for_slots<MG_XF / 8>([&](auto S_) { slot_group(S_, std::false_type{}); });
                } else {
                    mg_for_slots<MG_XF / 4>([&](auto S_) { slot_group(S_, std::true_type{}); });
                }
                if (d.MT == 1) {
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[0][e] += acc[1][e];
                }
                // ---- cross-warp reduction (K is split over the 8 warps) + epilogue by warps 0-3
                float* rb = red + (size_t)red_buf * (MG_WARPS * 2 * MG_RED_STRIDE) + (size_t)warp * (2 * MG_RED_STRIDE);
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    if (mt < d.MT) {
                        float* r2 = rb + mt * MG_RED_STRIDE;
                        r2[(2 * t) * 17 + g] = acc[mt][0];
                        r2[(2 * t + 1) * 17 + g] = acc[mt][1];
                        r2[(2 * t) * 17 + g + 8] = acc[mt][2];
                        r2[(2 * t + 1) * 17 + g + 8] = acc[mt][3];
                    }
                }
                if (dbg) { const long long t0 = clock64(); cta_sync(); sync_cyc += clock64() - t0; } else cta_sync();
                const long long te0 = dbg ? clock64() : 0;
                if (tid < 128) {
                    const int r = tid & 15, n = tid >> 4;
                    const float* r0 = red + (size_t)red_buf * (MG_WARPS * 2 * MG_RED_STRIDE) + n * 17 + r;
                    float s[2] = {0.f, 0.f};
#pragma unroll
                    for (int mt = 0; mt < 2; mt++)
                        if (mt < d.MT) {
#pragma unroll
                            for (int w = 0; w < MG_WARPS; w++) s[mt] += r0[(size_t)w * (2 * MG_RED_STRIDE) + mt * MG_RED_STRIDE];
                            float* pp = part + ((size_t)li * 2 + mt) * 128 + tid;
                            if (kb > 0) s[mt] += *pp;
                            if (!last_kb) *pp = s[mt];
                        }
                    if (last_kb) {
                        if (ph.kind == P3_MEGA_RESID) {                       // phi.py:483,485: h = bf16(h + bf16(y)); + sum of squares
                            bf16* out = reinterpret_cast<bf16*>(ph.out);
                            float sq = 0.f;
                            if (n < M) {
                                const bf16 hv = __float2bfloat16_rn(resid_pref + bf16_round(s[0]));
                                out[(size_t)n * ph.ldo + 16 * ti + r] = hv;
                                sq = __bfloat162float(hv) * __bfloat162float(hv);
                            }
#pragma unroll
                            for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                            if (r == 0 && ph.ss_out) ph.ss_out[(size_t)ti * 16 + n] = sq;
                        } else if (ph.kind == P3_MEGA_SWIGLU) {               // phi.py:470-471
                            if (n < M) {
                                const float gb = bf16_round(s[0]), ub = bf16_round(s[1]);
                                const float a = bf16_round(mg_silu(gb));
                                reinterpret_cast<bf16*>(ph.out)[(size_t)n * ph.ldo + 16 * ti + r] = __float2bfloat16_rn(a * ub);
                            }
                        } else if (ph.kind == P3_MEGA_F32) {                  // lm_head logits, phi.py:608
                            if (n < M) {
                                float* o = reinterpret_cast<float*>(ph.out) + (size_t)n * ph.ldo + 32 * ti + r;
                                o[0] = s[0]; o[16] = s[1];
                            }
                        } else if (n < M) {                                   // P3_MEGA_QKV_ROPE: phi.py:442-453, one new token per row
                            const int hd = P.hd, half = hd / 2, gpr = hd / 32, n_rope = (P.n_heads + P.n_kv) * gpr;
                            bf16* row = reinterpret_cast<bf16*>(ph.out) + (size_t)n * ph.ldo;
                            bf16* kd = reinterpret_cast<bf16*>(P.pool) + (size_t)kv_page * kv_page_elems(P.n_kv, hd) + (size_t)(past % P3_PAGE) * hd;
                            bf16* vd = kd + (size_t)P.n_kv * P3_PAGE * hd;
                            if (ti < n_rope) {
                                const int head = ti / gpr, q = ti % gpr, dd = q * 16 + r;
                                const float x1 = bf16_round(s[0]), x2 = bf16_round(s[1]);   // qkv_proj output is bf16 in the reference flow
                                float cs, sn;
                                if (gpr <= 3) {                                             // prefetched at kernel start (static register pick)
                                    cs = q == 0 ? rope_cs[0] : (q == 1 ? rope_cs[1] : rope_cs[2]);
                                    sn = q == 0 ? rope_sn[0] : (q == 1 ? rope_sn[1] : rope_sn[2]);
                                } else {
                                    const size_t tix = (size_t)n * P.tab_bstride + (size_t)past * half + dd;
                                    cs = P.cosT[tix]; sn = P.sinT[tix];
                                }
                                const bf16 o1 = __float2bfloat16_rn(x1 * cs - x2 * sn), o2 = __float2bfloat16_rn(x2 * cs + x1 * sn);
                                row[head * hd + dd] = o1;
                                row[head * hd + half + dd] = o2;
                                if (head >= P.n_heads) {
                                    bf16* k = kd + (size_t)(head - P.n_heads) * P3_PAGE * hd;
                                    k[dd] = o1; k[half + dd] = o2;
                                }
                            } else {
                                const int c0 = 32 * (ti - n_rope);
#pragma unroll
                                for (int mt = 0; mt < 2; mt++) {
                                    const bf16 v = __float2bfloat16_rn(s[mt]);
                                    const int c = c0 + 16 * mt + r;
                                    row[(P.n_heads + P.n_kv) * hd + c] = v;
                                    vd[(size_t)(c / hd) * P3_PAGE * hd + c % hd] = v;
                                }
                            }
                        }
                    }
                }
                if (dbg) epi_cyc += clock64() - te0;
                red_buf ^= 1;
            }
        }
        if (dbg && lane == 0) {
            if (warp == 0) { dbg[p * 8 + 3] = clock64(); dbg[p * 8 + 4] = wait_cyc; dbg[p * 8 + 6] = sync_cyc; dbg[p * 8 + 7] = epi_cyc; }
            if (warp == 7) dbg[p * 8 + 5] = wait_cyc;
        }
    }

    // ---- leave the barrier words clean for the next launch (every CTA has passed the last barrier by now)
    cta_sync();
    if (tid == 0 && gridDim.x > 1) {
        __threadfence();
        const unsigned old = atomicAdd(P.sync + 1, 1u);
        if (old == gridDim.x - 1) { P.sync[0] = 0u; P.sync[1] = 0u; __threadfence(); }
    }
}

extern "C" int p3_decode_mega_ctas(void) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return sms;
}

extern "C" int p3_decode_mega(const p3_mega_args* a, cudaStream_t st) {
    P3_CHECK_ARG(a && a->n_phases >= 1 && a->n_phases <= P3_MEGA_MAX_PHASES, "decode_mega: 1..%d phases", P3_MEGA_MAX_PHASES);
    P3_CHECK_ARG(a->M >= 1 && a->M <= 8, "decode_mega: M must be in [1,8] (got %d)", a->M);
    P3_CHECK_ARG(a->sync, "decode_mega: sync words are required");
    const int sms = p3_decode_mega_ctas();
    P3_CHECK_ARG(a->n_ctas >= 1 && a->n_ctas <= sms, "decode_mega: n_ctas %d must be in [1, #SM = %d] (all CTAs must be co-resident)", a->n_ctas, sms);
    for (int i = 0; i < a->n_phases; i++) {
        const p3_mega_phase& ph = a->ph[i];
        P3_CHECK_ARG(ph.kind >= P3_MEGA_RESID && ph.kind <= P3_MEGA_F32, "decode_mega: phase %d: unknown kind %d", i, ph.kind);
        P3_CHECK_ARG(mega_dims_ok(ph.kind, ph.N, ph.K, a->n_heads, a->n_kv, a->hd), "decode_mega: phase %d: unsupported shape N=%d K=%d", i, ph.N, ph.K);
        P3_CHECK_ARG(ph.wp && ph.x && ph.out && ph.cta_off && ph.tile_ids, "decode_mega: phase %d: null pointer", i);
        P3_CHECK_ARG(ph.ldx % 4 == 0, "decode_mega: phase %d: ldx must be a multiple of 4", i);
        P3_CHECK_ARG(ph.max_tiles_per_cta >= 0 && (ph.K <= MG_KBLOCK || ph.max_tiles_per_cta <= MG_MAX_PART),
                     "decode_mega: phase %d: at most %d tiles per CTA when K > %d", i, MG_MAX_PART, MG_KBLOCK);
        if (ph.kind == P3_MEGA_QKV_ROPE)
            P3_CHECK_ARG(a->cosT && a->sinT && a->pool && a->block_table, "decode_mega: QKV phase needs rope tables, pool and block table");
    }
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM);
        P3_CHECK_ARG(e == cudaSuccess, "decode_mega: smem attribute: %s", cudaGetErrorString(e));
        attr_set[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a->n_ctas); cfg.blockDim = dim3(MG_THREADS); cfg.dynamicSmemBytes = MG_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = p3_pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, decode_mega_kernel, *a);
    if (e != cudaSuccess) { p3_set_error("decode_mega: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}
