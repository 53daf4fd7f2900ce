// Skinny (decode-time) GEMM: Y[M,N] = X[M,K] . W[N,K]^T for M <= 16 tokens.
// Replaces nn.Linear qkv_proj / o_proj / gate_up_proj / down_proj / lm_head at decode
// (phi.py:437-438,465-466,604) where the op is a pure weight stream: HBM-bound, so the design
// goal is many 16-byte loads of W in flight per SM, not tensor-core occupancy.
//
// Mapping: W rows are the MMA "M" dimension (m16n8k16, bf16 -> fp32), the <=16 tokens are "N".
// Each lane owns 16 contiguous bytes of a W row per load; the k-index permutation this implies is
// applied identically to the X operand, so W needs no shared-memory transpose. A CTA owns 16*MT W
// rows and splits K over its 8 warps; partial sums are reduced through shared memory.
//
// Pipeline: every warp runs a private DEPTH-stage cp.async (LDGSTS) ring in shared memory that
// carries, per 64-wide k chunk, its W slices and the RMSNorm gains. Each lane reads back exactly the
// bytes it copied, so the ring needs no barrier (cp.async.wait_group only) and no registers: weight
// bytes in flight are bounded by shared memory (2 CTAs/SM x 8 warps x DEPTH stages = 192 KB/SM).
// The (L2-resident) activations are fetched into registers one chunk ahead.
//
// Fusions: RMSNorm prologue (phi.py:478-479: x*rsqrt(mean(x^2)+eps)*w, rounded to bf16),
// residual epilogue (phi.py:483,485), SwiGLU epilogue (phi.py:470-471), fp32 logits out.
#include "common.cuh"
#include "../../include/phi3_b200.h"

#define SK_WARPS 8
#define SK_THREADS (SK_WARPS * 32)
#define SK_SMEM_2CTA (112 * 1024)            // two CTAs per SM: 2 x (dynamic + 640 B static + 1 KB reserved) <= 228 KB

struct SkParams {
    const bf16* X; int64_t ldx;
    const bf16* norm_w; float eps;
    const bf16* W;
    const uint8_t* Wq; const bf16* Wmeta; // W4 variant: packed 4-bit codes [N][K/2] and (scale, bias) bf16 [N][K/64][2] (quant.py layout)
    void* out; int64_t ldo;
    const bf16* resid;
    int M, N, K, epi;
    const float* ss_in; int n_ss_in;     // per-token sum-of-squares partials [n_ss_in][16] written by the producer of X
    float* ss_out;                       // RESIDUAL epilogue: partial[blockIdx.x][16] of the rows this CTA produced
    // P3_EPI_ROPE_QKV: SuRoPE + paged-KV write fused into the qkv projection (phi.py:442-453)
    const float *cosT, *sinT; int64_t tab_bstride;
    int L, n_heads, n_kv, hd, past, row_div, write_cache;
    const int32_t* past_dev;
    bf16* pool; const int32_t* block_table; int bt_stride;
    const uint8_t* l2_pf; int64_t l2_pf_bytes;   // next kernel's weights, pulled into L2 while this one streams
    // RMSNorm split between producer and consumer (used by the 4-bit weight stream, where re-normalising X in every warp costs
    // more ALU than the stream leaves idle, and a separate norm kernel costs a launch): the RESIDUAL epilogue also writes
    // xg_out[m][n] = bf16(h[m][n] * xg_gain[n]) (gain of the NEXT norm); a consumer with rs_epi = 1 reads that as X and applies
    // rsqrt(mean(h^2) + eps) from ss_in to its reduced accumulators (norm_w must be null then).
    const bf16* xg_gain; bf16* xg_out; int64_t ldxg; int rs_epi;
    unsigned long long* trace;                   // tools/chain_trace.py (null in production)
    // 16-row tiles of W stored in stream order [tile][64-k chunk][row half][k half][lane][16 B] (model.pack_rows16): a stage is one
    // contiguous 2 KB block and a CTA reads one contiguous 16 x K block, instead of 8 rows x 64 B per instruction at a stride of
    // K x 2 bytes. Same bytes, same arithmetic; the memory round trip of a contiguous stream is about half (tools/sm_ingest_bench.cu),
    // which is what bounds the 192-CTA kernels (o_proj, down_proj: a CTA streams ring / round-trip bytes per us).
    int packed;
    int n_tiles;                                 // output tiles; the grid may be smaller (persistent CTAs, tile = blockIdx.x + i * gridDim.x)
};
#define P3_EPI_ROPE_QKV 7

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }

template <int NT, int MT, int DEPTH_, bool W4 = false>
struct SkCfg {
    static constexpr int W_SLOTS = 4 * MT, X_SLOTS = 0;                      // X travels in registers (L2-resident, 1 chunk ahead)
    static constexpr int CHUNK = W4 ? 128 : 64;                              // k elements per ring stage
    // bf16: W slices + 128 B of norm gains. W4: per (row tile, row half) 512 B of codes + 256 B of (scale, bias), + 256 B of gains
    static constexpr int STAGE = W4 ? (MT * 2 * 768 + 256) : (W_SLOTS * 512 + 128);
    static constexpr int META_OFF = MT * 2 * 512, GAIN_OFF = W4 ? MT * 2 * 768 : W_SLOTS * 512;
    static constexpr int DEPTH = DEPTH_;
    static constexpr int RING = SK_WARPS * DEPTH * STAGE;
    static constexpr int RED = SK_WARPS * MT * 8 * NT * 17 * 4;
    // ring + reduction scratch; the scratch is double buffered by tile parity when two CTAs per SM still fit with it
    static constexpr int NRED = (RING + 2 * RED <= SK_SMEM_2CTA) ? 2 : 1;
    static constexpr int SMEM = RING + NRED * RED;
};

// extracts the adjacent weight pair j of a packed word as bf16x2 (128 + q): nibbles j and j + 4 (quant.py::pack_w4g64)
__device__ __forceinline__ uint32_t w4_pair(uint32_t w, int j) { return ((w >> (4 * j)) & 0x000F000Fu) | 0x43004300u; }
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ float rs_from_parts(const float (*part)[16], int tok, int K, float eps) {
    float ss = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; w++) ss += part[w][tok];
    return rsqrtf(ss / (float)K + eps);
}

template <int NT, int MT, int DEPTH, bool W4>
__global__ void __launch_bounds__(SK_THREADS, (SkCfg<NT, MT, DEPTH, W4>::SMEM <= SK_SMEM_2CTA) ? 2 : 1) gemm_skinny_kernel(SkParams p) {
    using C = SkCfg<NT, MT, DEPTH, W4>;
    extern __shared__ __align__(128) uint8_t sk_smem[];
    __shared__ float s_rs[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int K = p.K, n_chunks = K / C::CHUNK;
    const uint32_t ring = smem_u32(sk_smem) + warp * (C::DEPTH * C::STAGE);

    // W rows of this lane (rows g and g+8 of each 16-row tile) for output tile `tile`; out_col0 / rope_head / rope_grp describe
    // the tile to the epilogue. A CTA walks tiles blockIdx.x, + gridDim.x, ... with ONE continuous weight ring: with a grid of
    // one wave (or less) no CTA ever starts a tile with an empty ring, and there is no second, partly filled wave.
    constexpr int PAIRS = (MT == 1) ? 8 : 16;                                   // (x1, x2) / (gate, up) pairs per tile
    auto tile_rows = [&](int tile, int (&wr)[MT][2], int& out_col0, int& rope_head, int& rope_grp) {
        rope_head = -1; rope_grp = 0;
        if (p.epi == P3_EPI_ROPE_QKV) {
            // tiles [0, (n_heads+n_kv)*hd/32): head h, group j -> tile 0 = cols h*hd + 16j.., tile 1 = the rotary
            // partners at +hd/2 (so each (x1,x2) pair meets in one CTA); remaining tiles: 32 V rows each
            // MT == 1: PAIRS = 8 dims per tile, x1 in rows g, the partners in rows g + 8 of the one 16-row tile; V tiles 16 rows
            const int gpr = p.hd / (2 * PAIRS), n_rope = (p.n_heads + p.n_kv) * gpr;
            int r0;
            if (tile < n_rope) {
                rope_head = tile / gpr; rope_grp = tile % gpr;
                r0 = rope_head * p.hd + rope_grp * PAIRS;
                out_col0 = r0;
                if (MT == 1) {
                    wr[0][0] = r0 + g; wr[0][1] = r0 + p.hd / 2 + g;
                } else {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        wr[mt][0] = r0 + mt * (p.hd / 2) + g;
                        wr[mt][1] = r0 + mt * (p.hd / 2) + g + 8;
                    }
                }
            } else {
                r0 = (p.n_heads + p.n_kv) * p.hd + (tile - n_rope) * 16 * MT;
                out_col0 = r0;
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    wr[mt][0] = r0 + mt * 16 + g;
                    wr[mt][1] = r0 + mt * 16 + g + 8;
                }
            }
        } else if (p.epi == P3_EPI_SWIGLU) {
            // interleaved gate/up layout: [128 gate rows | 128 up rows] per 256-row block
            int o0 = tile * PAIRS;
            int gate0 = (o0 / 128) * 256 + (o0 % 128);
            out_col0 = o0;
            if (MT == 1) {                                                       // 8 gate rows (g) + their 8 up rows (g + 8)
                wr[0][0] = gate0 + g; wr[0][1] = gate0 + 128 + g;
            } else {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    wr[mt][0] = gate0 + mt * 128 + g;
                    wr[mt][1] = gate0 + mt * 128 + g + 8;
                }
            }
        } else {
            int n0 = tile * 16 * MT;
            out_col0 = n0;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                wr[mt][0] = min(n0 + mt * 16 + g, p.N - 1);
                wr[mt][1] = min(n0 + mt * 16 + g + 8, p.N - 1);
            }
        }
    };
    const int n_tiles = p.n_tiles;
    const bf16* wrow[MT][2];                                                     // bf16 stream: 16 B (8 weights) per lane and load
    const uint8_t* qrow[MT][2]; const bf16* mrow[MT][2];                         // W4 stream: 16 B (32 weights) + 8 B (2 groups' scale, bias)
    auto issue_rows = [&](int tile) {                                            // row pointers of the ISSUE cursor's tile
        if (MT == 1 && p.packed) { wrow[0][0] = p.W + (size_t)tile * 16 * K + lane * 8; return; }
        int wr[MT][2], oc, rh, rg;
        tile_rows(tile, wr, oc, rh, rg);
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                wrow[mt][hh] = W4 ? nullptr : p.W + (size_t)wr[mt][hh] * K + t * 8;
                qrow[mt][hh] = W4 ? p.Wq + (size_t)wr[mt][hh] * (K / 2) + t * 16 : nullptr;
                mrow[mt][hh] = W4 ? p.Wmeta + (size_t)wr[mt][hh] * (K / 64) * 2 : nullptr;
            }
    };
    const bf16* xrow[NT];
    int xbytes[NT];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        int m = nt * 8 + g;
        xbytes[nt] = m < p.M ? 16 : 0;                                          // rows >= M are zero-filled
        xrow[nt] = p.X + (size_t)(m < p.M ? m : 0) * p.ldx + t * (W4 ? 16 : 8);
    }

    // stage layout (per warp): W slots [mt][row-half][k-half], X slots [nt][k-half], each 32 lanes x 16 B;
    // then 128 B of norm gains (64 elements, read back with broadcast)
    const uint64_t pol = l2_evict_first_policy();
    auto issue_w = [&](int ci, int stage) {                                      // weights + norm gains: immutable
        if constexpr (W4) {
            const uint32_t sb = ring + stage * C::STAGE;
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    cp_async16_stream(sb + (mt * 2 + hh) * 512 + lane * 16, qrow[mt][hh] + ci * 64, pol);
                    cp_async8(sb + C::META_OFF + (mt * 2 + hh) * 256 + lane * 8, mrow[mt][hh] + ci * 4);
                }
            if (p.norm_w && lane < 16) cp_async16(sb + C::GAIN_OFF + lane * 16, p.norm_w + ci * 128 + lane * 8);
            return;
        }
        const uint32_t sb = ring + stage * C::STAGE + lane * 16;
        const int k0 = ci * 64;
        if (MT == 1 && p.packed) {
            const bf16* src = wrow[0][0] + (size_t)ci * 1024;
#pragma unroll
            for (int sl = 0; sl < 4; sl++) cp_async16_stream(sb + sl * 512, src + sl * 256, pol);
        } else
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                cp_async16_stream(sb + ((mt * 2 + hh) * 2 + 0) * 512, wrow[mt][hh] + k0, pol);
                cp_async16_stream(sb + ((mt * 2 + hh) * 2 + 1) * 512, wrow[mt][hh] + k0 + 32, pol);
            }
        if (p.norm_w && lane < 8)
            cp_async16(ring + stage * C::STAGE + (C::W_SLOTS + C::X_SLOTS) * 512 + lane * 16, p.norm_w + k0 + lane * 8);
    };
    uint4 xnext[NT][W4 ? 4 : 2];
    auto load_x = [&](int ci) {                                                  // activations: produced by the previous kernel
        if constexpr (W4) {                                                      // lane t: k = 128 ci + 64 grp + 16 t + (0..15)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int q = 0; q < 4; q++)
                    xnext[nt][q] = xbytes[nt] ? *reinterpret_cast<const uint4*>(xrow[nt] + ci * 128 + (q >> 1) * 64 + (q & 1) * 8)
                                              : make_uint4(0, 0, 0, 0);
            return;
        }
        const int k0 = ci * 64;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            if (xbytes[nt]) {
                xnext[nt][0] = *reinterpret_cast<const uint4*>(xrow[nt] + k0);
                xnext[nt][1] = *reinterpret_cast<const uint4*>(xrow[nt] + k0 + 32);
            } else {
                xnext[nt][0] = make_uint4(0, 0, 0, 0); xnext[nt][1] = make_uint4(0, 0, 0, 0);
            }
        }
    };

    // Fill the ring with weights first: they do not depend on the previous kernel, so under
    // programmatic dependent launch this HBM traffic overlaps the predecessor's tail.
    trace_stamp(p.trace, blockIdx.x, 0);
    trace_meta(p.trace, blockIdx.x, p.epi * 100 + MT * 10 + (W4 ? 1 : 0));
    pdl_trigger();
    // issue cursor: (tile, chunk) of the next weight stage to request; runs DEPTH stages ahead of the consumer, across tiles
    int itile = (warp < n_chunks) ? (int)blockIdx.x : n_tiles, ici = warp;
    if (itile < n_tiles) issue_rows(itile);
    auto issue_next = [&](int stage) {
        if (itile < n_tiles) {
            issue_w(ici, stage);
            ici += SK_WARPS;
            if (ici >= n_chunks) {
                ici = warp; itile += gridDim.x;
                if (itile < n_tiles) issue_rows(itile);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < C::DEPTH; s++) issue_next(s);
    if (p.l2_pf) l2_prefetch_slice(p.l2_pf, p.l2_pf_bytes, blockIdx.x, gridDim.x, tid, SK_THREADS);
    pdl_wait();
    trace_stamp(p.trace, blockIdx.x, 1);
    if (warp < n_chunks) load_x(warp);

    // ---- RMSNorm prologue: rs[m] = rsqrt(mean(x^2) + eps), from the producer's partial sums when
    // available (fixed summation order: deterministic), else recomputed from X (L2 resident)
    // rs_epi: the scale is needed only in the epilogue -> request the partial sums now, reduce them after the K loop (reduced
    // here they would sit on every consumer's critical path). Coalesced: a warp instruction reads 8 partial rows x 16 tokens
    // as float4 (lane = 4 * row + token quad); the [c * 16 + m] gather of the fused-norm prologue costs 32 sectors per
    // instruction, ~1.6 us of LSU time per CTA, which the short 4-bit kernels cannot hide.
    constexpr int SSR = 4;
    __shared__ float s_sspart[SK_WARPS][16];
    float4 ss_raw[SSR];
    const bool ss_late = p.rs_epi && p.n_ss_in <= 64 * SSR;
    if (ss_late) {
#pragma unroll
        for (int j = 0; j < SSR; j++) {
            const int c = j * 64 + warp * 8 + (lane >> 2);
            ss_raw[j] = (c < p.n_ss_in) ? __ldcg(reinterpret_cast<const float4*>(p.ss_in + c * 16) + (lane & 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else if ((p.norm_w || p.rs_epi) && p.ss_in) {                              // same coalesced read, reduced now (X is scaled in the loop)
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = warp * 8 + (lane >> 2); c < p.n_ss_in; c += 8 * SK_WARPS) {
            const float4 f = __ldcg(reinterpret_cast<const float4*>(p.ss_in + c * 16) + (lane & 3));
            a.x += f.x; a.y += f.y; a.z += f.z; a.w += f.w;
        }
#pragma unroll
        for (int sh = 4; sh < 32; sh <<= 1) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, sh); a.y += __shfl_xor_sync(0xffffffffu, a.y, sh);
            a.z += __shfl_xor_sync(0xffffffffu, a.z, sh); a.w += __shfl_xor_sync(0xffffffffu, a.w, sh);
        }
        if (lane < 4) *reinterpret_cast<float4*>(&s_sspart[warp][4 * lane]) = a;
        __syncthreads();
        if (tid < 16) s_rs[tid] = rs_from_parts(s_sspart, tid, K, p.eps);
        __syncthreads();
    } else if (p.norm_w) {
        for (int m = warp; m < p.M; m += SK_WARPS) {
            const uint4* xr = reinterpret_cast<const uint4*>(p.X + (size_t)m * p.ldx);
            float ss = 0.f;
            for (int c = lane; c < K / 8; c += 32) {
                uint4 v = xr[c];
                const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
                for (int j = 0; j < 4; j++) { float2 f = unpack_bf16(u[j]); ss += f.x * f.x + f.y * f.y; }
            }
            ss = warp_sum(ss);
            if (lane == 0) s_rs[m] = rsqrtf(ss / (float)K + p.eps);
        }
        __syncthreads();
    }

    float rs[NT];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) rs[nt] = (p.norm_w && nt * 8 + g < p.M) ? s_rs[nt * 8 + g] : 1.f;
    int stage = 0;
    bool first_tile = true;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first_tile = false) {
    int wr_unused[MT][2], out_col0, rope_head, rope_grp;
    tile_rows(tile, wr_unused, out_col0, rope_head, rope_grp);
    // ROPE_QKV epilogue: this thread's rope factors and page id are known up front -> request them now, use them after the K
    // loop (they can miss L2; fetched in the epilogue they would sit on the kernel's tail)
    float rope_cs = 1.f, rope_sn = 0.f;
    int rope_page = 0;
    if (p.epi == P3_EPI_ROPE_QKV && tid < 8 * NT * PAIRS && (tid / PAIRS) < p.M) {
        const int tok = tid / PAIRS, r = tid % PAIRS;
        const int past0 = p.past_dev ? *p.past_dev : p.past;
        const int b = tok / p.L, pos = past0 + tok % p.L;
        if (rope_head >= 0) {
            const size_t ti = (size_t)(b / p.row_div) * p.tab_bstride + (size_t)pos * (p.hd / 2) + rope_grp * PAIRS + r;
            rope_cs = __ldg(p.cosT + ti); rope_sn = __ldg(p.sinT + ti);
        }
        if (p.write_cache) rope_page = __ldg(p.block_table + (size_t)(b / p.row_div) * p.bt_stride + pos / P3_PAGE);
    }
    // RESIDUAL epilogue (MT == 1): this thread's output element is known up front -> fetch the residual now
    float resid_pref = 0.f, gain_pref = 1.f;
    if (p.epi == P3_EPI_RESIDUAL && MT == 1 && tid < 8 * NT * 16) {
        int r = tid & 15, tok = tid >> 4, n = out_col0 + r;
        if (tok < p.M && n < p.N) {
            resid_pref = __bfloat162float(p.resid[(size_t)tok * p.ldo + n]);
            if (p.xg_out) gain_pref = __bfloat162float(p.xg_gain[n]);             // immutable: could even precede pdl_wait
        }
    }
    float acc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[mt][nt][j] = 0.f;

    for (int ci = warp; ci < n_chunks; ci += SK_WARPS) {
        cp_async_wait<C::DEPTH - 1>();                                           // oldest group (this chunk) has landed
        __syncwarp();                                                            // norm gains were copied by lanes 0-7
        if (ci == warp && first_tile) trace_stamp(p.trace, blockIdx.x, 2);
        const uint8_t* sb = sk_smem + (size_t)warp * (C::DEPTH * C::STAGE) + stage * C::STAGE;
        if constexpr (W4) {
            uint4 qc[MT][2]; uint2 qm[MT][2]; uint4 xq[NT][4];
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    qc[mt][hh] = *reinterpret_cast<const uint4*>(sb + (mt * 2 + hh) * 512 + lane * 16);
                    qm[mt][hh] = *reinterpret_cast<const uint2*>(sb + C::META_OFF + (mt * 2 + hh) * 256 + lane * 8);
                }
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int q = 0; q < 4; q++) xq[nt][q] = xnext[nt][q];
            load_x(ci + SK_WARPS < n_chunks ? ci + SK_WARPS : warp);          // next chunk's X (wraps into the next tile)
            if (p.norm_w) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint4 nw = *reinterpret_cast<const uint4*>(sb + C::GAIN_OFF + ((q >> 1) * 64 + t * 16 + (q & 1) * 8) * 2);
                    const uint32_t* uw = reinterpret_cast<const uint32_t*>(&nw);
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        uint32_t* ux = reinterpret_cast<uint32_t*>(&xq[nt][q]);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            float2 f = unpack_bf16(ux[j]), w = unpack_bf16(uw[j]);
                            ux[j] = pack_bf16(f.x * rs[nt] * w.x, f.y * rs[nt] * w.y);
                        }
                    }
                }
            }
            __syncwarp();
            issue_next(stage);
            if (++stage == C::DEPTH) stage = 0;
            // per 64-wide group: MMA on the raw codes (as bf16 128+q) and on a matrix of ones (-> sum of x), then
            // y += scale * sum((128+q) x) + (bias - 128 scale) * sum(x)   ==  sum((scale q + bias) x)  in fp32
            const uint32_t ones[4] = {0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u};
#pragma unroll
            for (int gi = 0; gi < 2; gi++) {
                float accg[MT][NT][4], accx[NT][4];
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
#pragma unroll
                    for (int j = 0; j < 4; j++) accx[nt][j] = 0.f;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int j = 0; j < 4; j++) accg[mt][nt][j] = 0.f;
                }
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const int i = s >> 1, j0 = 2 * (s & 1);
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        const uint32_t* ux = reinterpret_cast<const uint32_t*>(&xq[nt][gi * 2 + i]);
                        mma_bf16_16816(accx[nt], ones, ux[j0], ux[j0 + 1]);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const uint32_t* c0 = reinterpret_cast<const uint32_t*>(&qc[mt][0]);
                        const uint32_t* c1 = reinterpret_cast<const uint32_t*>(&qc[mt][1]);
                        const uint32_t w0 = c0[gi * 2 + i], w1 = c1[gi * 2 + i];
                        uint32_t a[4] = {w4_pair(w0, j0), w4_pair(w1, j0), w4_pair(w0, j0 + 1), w4_pair(w1, j0 + 1)};
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            const uint32_t* ux = reinterpret_cast<const uint32_t*>(&xq[nt][gi * 2 + i]);
                            mma_bf16_16816(accg[mt][nt], a, ux[j0], ux[j0 + 1]);
                        }
                    }
                }
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const float2 m0 = unpack_bf16(gi ? qm[mt][0].y : qm[mt][0].x);   // (scale, bias) of row g
                    const float2 m1 = unpack_bf16(gi ? qm[mt][1].y : qm[mt][1].x);   // row g + 8
                    const float c0 = m0.y - 128.f * m0.x, c1 = m1.y - 128.f * m1.x;
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        acc[mt][nt][0] += m0.x * accg[mt][nt][0] + c0 * accx[nt][0];
                        acc[mt][nt][1] += m0.x * accg[mt][nt][1] + c0 * accx[nt][1];
                        acc[mt][nt][2] += m1.x * accg[mt][nt][2] + c1 * accx[nt][2];
                        acc[mt][nt][3] += m1.x * accg[mt][nt][3] + c1 * accx[nt][3];
                    }
                }
            }
            continue;
        }
        uint4 wcur[MT][2][2], xf[NT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int hh = 0; hh < 2; hh++)
#pragma unroll
                for (int kh = 0; kh < 2; kh++)
                    wcur[mt][hh][kh] = *reinterpret_cast<const uint4*>(sb + ((mt * 2 + hh) * 2 + kh) * 512 + lane * 16);
#pragma unroll
        for (int nt = 0; nt < NT; nt++) { xf[nt][0] = xnext[nt][0]; xf[nt][1] = xnext[nt][1]; }
        load_x(ci + SK_WARPS < n_chunks ? ci + SK_WARPS : warp);                  // next chunk's X, one iteration ahead (wraps into the next tile)
        if (p.norm_w) {
            uint4 nw[2];
            nw[0] = *reinterpret_cast<const uint4*>(sb + (C::W_SLOTS + C::X_SLOTS) * 512 + t * 16);
            nw[1] = *reinterpret_cast<const uint4*>(sb + (C::W_SLOTS + C::X_SLOTS) * 512 + 64 + t * 16);
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int kh = 0; kh < 2; kh++) {
                    uint32_t* ux = reinterpret_cast<uint32_t*>(&xf[nt][kh]);
                    const uint32_t* uw = reinterpret_cast<const uint32_t*>(&nw[kh]);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float2 f = unpack_bf16(ux[j]), w = unpack_bf16(uw[j]);
                        ux[j] = pack_bf16(f.x * rs[nt] * w.x, f.y * rs[nt] * w.y);
                    }
                }
        }
        // the stage is in registers now: refill it with the chunk DEPTH iterations ahead
        __syncwarp();
        issue_next(stage);
        if (++stage == C::DEPTH) stage = 0;
#pragma unroll
        for (int kh = 0; kh < 2; kh++)
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const uint32_t* w0 = reinterpret_cast<const uint32_t*>(&wcur[mt][0][kh]);
                    const uint32_t* w1 = reinterpret_cast<const uint32_t*>(&wcur[mt][1][kh]);
                    uint32_t a[4] = {w0[2 * s], w1[2 * s], w0[2 * s + 1], w1[2 * s + 1]};
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        const uint32_t* ux = reinterpret_cast<const uint32_t*>(&xf[nt][kh]);
                        mma_bf16_16816(acc[mt][nt], a, ux[2 * s], ux[2 * s + 1]);
                    }
                }
    }
    if (ss_late && first_tile) {                                                 // fixed order everywhere: deterministic
        float4 a = ss_raw[0];
#pragma unroll
        for (int j = 1; j < SSR; j++) { a.x += ss_raw[j].x; a.y += ss_raw[j].y; a.z += ss_raw[j].z; a.w += ss_raw[j].w; }
#pragma unroll
        for (int sh = 4; sh < 32; sh <<= 1) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, sh); a.y += __shfl_xor_sync(0xffffffffu, a.y, sh);
            a.z += __shfl_xor_sync(0xffffffffu, a.z, sh); a.w += __shfl_xor_sync(0xffffffffu, a.w, sh);
        }
        if (lane < 4) *reinterpret_cast<float4*>(&s_sspart[warp][4 * lane]) = a;
    }
    if (tile + (int)gridDim.x >= n_tiles) {                                      // this CTA's stream is over: its share of HBM is idle from here on
        trace_stamp(p.trace, blockIdx.x, 3);
    }

    // ---- cross-warp reduction through shared memory. The scratch sits behind the ring and alternates between two buffers, so
    // one barrier per tile is enough: a warp can only write buffer (i + 2) & 1 after the barrier of tile i + 1, which every warp
    // reaches after it has finished reading buffer i & 1 in the epilogue of tile i. The ring keeps streaming meanwhile.
    float (*s_red)[MT][8 * NT][17] = reinterpret_cast<float (*)[MT][8 * NT][17]>(
        sk_smem + C::RING + (C::NRED == 2 ? (((tile - (int)blockIdx.x) / (int)gridDim.x) & 1) * C::RED : 0));
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            s_red[warp][mt][nt * 8 + 2 * t][g] = acc[mt][nt][0];
            s_red[warp][mt][nt * 8 + 2 * t + 1][g] = acc[mt][nt][1];
            s_red[warp][mt][nt * 8 + 2 * t][g + 8] = acc[mt][nt][2];
            s_red[warp][mt][nt * 8 + 2 * t + 1][g + 8] = acc[mt][nt][3];
        }
    __syncthreads();

    if (p.epi == P3_EPI_ROPE_QKV) {
        const int past = p.past_dev ? *p.past_dev : p.past;
        const int half = p.hd / 2, qkv_dim = (p.n_heads + 2 * p.n_kv) * p.hd;
        bf16* outp = reinterpret_cast<bf16*>(p.out);
        for (int o = tid; o < 8 * NT * PAIRS; o += SK_THREADS) {
            int r = o % PAIRS, tok = o / PAIRS;
            if (tok >= p.M) continue;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; w++) { a0 += s_red[w][0][tok][r]; a1 += (MT == 1) ? s_red[w][0][tok][r + 8] : s_red[w][MT - 1][tok][r]; }
            if (p.rs_epi) { const float rsv = ss_late ? rs_from_parts(s_sspart, tok, K, p.eps) : s_rs[tok]; a0 *= rsv; a1 *= rsv; }
            const int b = tok / p.L, pos = past + tok % p.L;
            bf16* row = outp + (size_t)tok * p.ldo;
            bf16 *kd = nullptr, *vd = nullptr;
            if (p.write_cache) {
                const int page = (o == tid) ? rope_page : p.block_table[(size_t)(b / p.row_div) * p.bt_stride + pos / P3_PAGE];
                kd = p.pool + (size_t)page * kv_page_elems(p.n_kv, p.hd) + (size_t)(pos % P3_PAGE) * p.hd;
                vd = kd + (size_t)p.n_kv * P3_PAGE * p.hd;
            }
            if (rope_head >= 0) {                                                  // q or k head: rotate (phi.py:418-423)
                const float x1 = bf16_round(a0), x2 = bf16_round(a1);              // qkv_proj output is bf16 in the reference flow
                const int d = rope_grp * PAIRS + r;
                const size_t ti = (size_t)(b / p.row_div) * p.tab_bstride + (size_t)pos * half + d;
                const float cs = (o == tid) ? rope_cs : p.cosT[ti], sn = (o == tid) ? rope_sn : p.sinT[ti];
                const bf16 o1 = __float2bfloat16_rn(x1 * cs - x2 * sn), o2 = __float2bfloat16_rn(x2 * cs + x1 * sn);
                row[rope_head * p.hd + d] = o1;
                row[rope_head * p.hd + half + d] = o2;
                if (kd && rope_head >= p.n_heads) {
                    bf16* k = kd + (size_t)(rope_head - p.n_heads) * P3_PAGE * p.hd;
                    k[d] = o1; k[half + d] = o2;
                }
            } else {                                                               // v rows: copy to qkv buffer + cache
                const int c0 = out_col0 - (p.n_heads + p.n_kv) * p.hd;             // column inside the V block
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    const bf16 v = __float2bfloat16_rn(mt == 0 ? a0 : a1);
                    const int c = c0 + mt * PAIRS + r;
                    row[(p.n_heads + p.n_kv) * p.hd + c] = v;
                    if (vd) vd[(size_t)(c / p.hd) * P3_PAGE * p.hd + c % p.hd] = v;
                }
            }
        }
        (void)qkv_dim;
    } else if (p.epi == P3_EPI_SWIGLU) {
        for (int o = tid; o < 8 * NT * PAIRS; o += SK_THREADS) {
            int r = o % PAIRS, tok = o / PAIRS;
            if (tok >= p.M) continue;
            float gsum = 0.f, usum = 0.f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; w++) { gsum += s_red[w][0][tok][r]; usum += (MT == 1) ? s_red[w][0][tok][r + 8] : s_red[w][MT - 1][tok][r]; }
            if (p.rs_epi) { const float rsv = ss_late ? rs_from_parts(s_sspart, tok, K, p.eps) : s_rs[tok]; gsum *= rsv; usum *= rsv; }
            float gb = bf16_round(gsum), ub = bf16_round(usum);
            float a = bf16_round(silu_f(gb));
            reinterpret_cast<bf16*>(p.out)[(size_t)tok * p.ldo + out_col0 + r] = __float2bfloat16_rn(a * ub);
        }
    } else {
        for (int o = tid; o < MT * 8 * NT * 16; o += SK_THREADS) {                  // whole warps enter (128|256|512 outputs)
            int r = o & 15, mt = (o >> 4) % MT, tok = o / (16 * MT);
            int n = out_col0 + mt * 16 + r;
            const bool ok = tok < p.M && n < p.N;
            float s = 0.f, sq = 0.f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; w++) s += s_red[w][mt][tok][r];
            if (p.rs_epi && tok < 16) s *= ss_late ? rs_from_parts(s_sspart, tok, K, p.eps) : s_rs[tok];
            size_t off = (size_t)tok * p.ldo + n;
            if (ok) {
                if (p.epi == P3_EPI_F32) {
                    reinterpret_cast<float*>(p.out)[off] = s;
                } else if (p.epi == P3_EPI_RESIDUAL) {
                    float rv = (MT == 1) ? resid_pref : __bfloat162float(p.resid[off]);
                    bf16 hv = __float2bfloat16_rn(rv + bf16_round(s));
                    reinterpret_cast<bf16*>(p.out)[off] = hv;
                    sq = __bfloat162float(hv) * __bfloat162float(hv);
                    if (p.xg_out) p.xg_out[(size_t)tok * p.ldxg + n] = __float2bfloat16_rn(__bfloat162float(hv) * ((MT == 1) ? gain_pref : __bfloat162float(p.xg_gain[n])));
                } else {
                    reinterpret_cast<bf16*>(p.out)[off] = __float2bfloat16_rn(s);
                }
            }
            if (p.ss_out && MT == 1) {                                             // 16 lanes = one token's 16 new columns
#pragma unroll
                for (int d = 8; d > 0; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
                if (r == 0 && tok < 16) p.ss_out[(size_t)tile * 16 + tok] = (tok < p.M) ? sq : 0.f;
            }
        }
    }
    if (C::NRED == 1 && tile + (int)gridDim.x < n_tiles) __syncthreads();        // single scratch buffer: drained before the next tile's sums
    }   // tile loop
    trace_stamp(p.trace, blockIdx.x, 4);
}

template <int NT, int MT, int DEPTH, bool W4 = false>
static int launch_skinny_d(const SkParams& p, unsigned grid, cudaStream_t st) {
    using C = SkCfg<NT, MT, DEPTH, W4>;
    static P3DevFlags flags; bool& set = flags.cur();
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_skinny_kernel<NT, MT, DEPTH, W4>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        P3_CHECK_ARG(e == cudaSuccess, "gemm_skinny: smem attribute: %s", cudaGetErrorString(e));
        set = true;
    }
    p3_launch_pdl(gemm_skinny_kernel<NT, MT, DEPTH, W4>, dim3(grid), dim3(SK_THREADS), (size_t)C::SMEM, st, p);
    P3_CHECK_LAUNCH("gemm_skinny");
    return 0;
}

// ring depth: default keeps 2 CTAs/SM; P3_SK_DEPTH overrides (tuning only)
static int sk_depth_override() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("P3_SK_DEPTH"); v = e ? atoi(e) : 0; }
    return v;
}
// pair tiles (qkv + rope, gate_up + SwiGLU) as ONE 16-row MMA tile (8 + 8 rows, the 4-stage ring of the plain kernels) instead of two
static int sk_pair_mt1() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("P3_SK_MT1"); v = e ? atoi(e) : 0; }
    return v;
}
// Persistent grid: at most one wave of CTAs (2 per SM) walks the output tiles. In-kernel stamps (profiles/r02_chain_trace_*.log):
// with one CTA per tile gate_up (512 tiles) ran a second, partly filled wave whose CTAs all started with an empty ring: 20.1 us
// for a 15.6 us stream; one wave with a continuous ring: 19.4 us, decode step -2 % (profiles/r02_skinny_persistent_ab.log).
// P3_SK_GRID=0 restores one CTA per tile, any other value is the cap (1 CTA/SM = 148 measured 25 % slower: half the bytes in flight).
static int sk_grid_cap() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("P3_SK_GRID");
        if (e) v = atoi(e);
        else { int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); v = 2 * sms; }
    }
    return v;
}
template <int NT, int MT>
static int launch_skinny(const SkParams& p_, unsigned grid, cudaStream_t st) {
    SkParams p = p_; p.trace = p3_trace_slot();
    p.n_tiles = (int)grid;
    const int cap = sk_grid_cap();
    if (cap > 0 && grid > (unsigned)cap) {
        // every CTA gets the same number of tiles (gate_up: 256 CTAs x 2 instead of 216 x 2 + 80 x 1: -0.9 % per decode step;
        // the slots left free take the next kernel's first CTAs, which fill their rings early). P3_SK_EVEN=0: fill the cap.
        static int even = -1;
        if (even < 0) { const char* e = getenv("P3_SK_EVEN"); even = e ? atoi(e) : 1; }
        const unsigned per = (grid + cap - 1) / cap;
        grid = even ? (grid + per - 1) / per : (unsigned)cap;
    }
    // 4-bit stream: 4 stages x 128 k per warp; 3 for the 32-row tiles so that ring + reduction scratch keep two CTAs per SM
    // (4 stages = 115 KB = one CTA per SM: 2.83 instead of 2.55 ms per decode step)
    if (p.Wq) return launch_skinny_d<NT, MT, (MT == 2 ? 3 : 4), true>(p, grid, st);
    int d = sk_depth_override();
    { static int d1 = -1; if (d1 < 0) { const char* e = getenv("P3_SK_DEPTH1"); d1 = e ? atoi(e) : 0; } if (MT == 1 && d1) d = d1; }
    if (d == 0) d = (MT == 1) ? 4 : 2;                       // measured best (tools/microbench.py): deeper rings do not pay, residency does
    switch (d) {
        case 2: return launch_skinny_d<NT, MT, 2>(p, grid, st);
        case 3: return launch_skinny_d<NT, MT, 3>(p, grid, st);
        case 4: return launch_skinny_d<NT, MT, 4>(p, grid, st);
        case 6: return launch_skinny_d<NT, MT, 6>(p, grid, st);
        case 8: return launch_skinny_d<NT, MT, (MT == 1 ? 8 : 4)>(p, grid, st);
        default: return launch_skinny_d<NT, MT, 2>(p, grid, st);
    }
}

static int skinny_impl(const void* X, int64_t ldx, const void* norm_w, float eps, const void* W, const void* Wq,
                       const void* Wmeta, void* out, int64_t ldo, const void* resid, int M, int N, int K, int epi,
                       const float* ss_in, int n_ss_in, float* ss_out, const void* l2_prefetch, int64_t l2_prefetch_bytes,
                       cudaStream_t st, const void* xg_gain = nullptr, void* xg_out = nullptr, int64_t ldxg = 0, int rs_epi = 0,
                       int packed = 0) {
    P3_CHECK_ARG(M >= 1 && M <= 16, "gemm_skinny: M must be in [1,16] (got %d)", M);
    P3_CHECK_ARG(!packed || (W && !Wq && N % 16 == 0 && N < 148 * 32 * 2 && epi != P3_EPI_SWIGLU),
                 "gemm_skinny: the packed layout is for bf16 matrices streamed as 16-row tiles (N %% 16 == 0, N < 9472, no SwiGLU)");
    P3_CHECK_ARG(!rs_epi || (ss_in && !norm_w), "gemm_skinny: rs_epi needs ss_in and no norm_w (X is already gain-scaled)");
    P3_CHECK_ARG(!xg_out || (xg_gain && epi == P3_EPI_RESIDUAL), "gemm_skinny: xg_out needs xg_gain and the residual epilogue");
    P3_CHECK_ARG(K % 64 == 0, "gemm_skinny: K must be a multiple of 64 (got %d)", K);
    P3_CHECK_ARG(!Wq || (K % 128 == 0 && Wmeta), "gemm_skinny_w4: K must be a multiple of 128 and meta must be given");
    P3_CHECK_ARG(epi == P3_EPI_NONE || epi == P3_EPI_RESIDUAL || epi == P3_EPI_SWIGLU || epi == P3_EPI_F32,
                 "gemm_skinny: unsupported epilogue %d", epi);
    P3_CHECK_ARG(epi != P3_EPI_RESIDUAL || resid, "gemm_skinny: residual epilogue needs resid");
    P3_CHECK_ARG(ldx % 8 == 0, "gemm_skinny: ldx must be a multiple of 8");
    P3_CHECK_ARG(!ss_out || (epi == P3_EPI_RESIDUAL && N < 148 * 32 * 2), "gemm_skinny: ss_out needs the 16-row residual configuration");
    SkParams p{};
    p.X = (const bf16*)X; p.ldx = ldx; p.norm_w = (const bf16*)norm_w; p.eps = eps; p.W = (const bf16*)W; p.out = out;
    p.Wq = (const uint8_t*)Wq; p.Wmeta = (const bf16*)Wmeta;
    p.ldo = ldo; p.resid = (const bf16*)resid; p.M = M; p.N = N; p.K = K; p.epi = epi;
    p.ss_in = ss_in; p.n_ss_in = n_ss_in; p.ss_out = ss_out;
    p.l2_pf = (const uint8_t*)l2_prefetch; p.l2_pf_bytes = l2_prefetch_bytes;
    p.xg_gain = (const bf16*)xg_gain; p.xg_out = (bf16*)xg_out; p.ldxg = ldxg; p.rs_epi = rs_epi; p.packed = packed;
    if (epi == P3_EPI_SWIGLU) {
        P3_CHECK_ARG(N % 256 == 0, "gemm_skinny: SwiGLU needs N (gate+up rows) to be a multiple of 256");
        if (sk_pair_mt1() && M <= 8) return launch_skinny<1, 1>(p, (unsigned)(N / 2 / 8), st);
        unsigned grid = (unsigned)(N / 2 / 16);
        return M <= 8 ? launch_skinny<1, 2>(p, grid, st) : launch_skinny<2, 2>(p, grid, st);
    }
    if (N >= 148 * 32 * 2) {
        unsigned grid = (unsigned)((N + 31) / 32);
        return M <= 8 ? launch_skinny<1, 2>(p, grid, st) : launch_skinny<2, 2>(p, grid, st);
    }
    unsigned grid = (unsigned)((N + 15) / 16);
    return M <= 8 ? launch_skinny<1, 1>(p, grid, st) : launch_skinny<2, 1>(p, grid, st);
}

extern "C" int p3_gemm_skinny(const void* X, int64_t ldx, const void* norm_w, float eps, const void* W, void* out,
                              int64_t ldo, const void* resid, int M, int N, int K, int epi, const float* ss_in,
                              int n_ss_in, float* ss_out, const void* l2_prefetch, int64_t l2_prefetch_bytes,
                              cudaStream_t st) {
    return skinny_impl(X, ldx, norm_w, eps, W, nullptr, nullptr, out, ldo, resid, M, N, K, epi, ss_in, n_ss_in, ss_out,
                       l2_prefetch, l2_prefetch_bytes, st);
}

// quantize_model=True (pv:264,291-305): the same op over the 4-bit g64 image of W (quant.py::pack_w4g64 layout)
extern "C" int p3_gemm_skinny_w4(const void* X, int64_t ldx, const void* norm_w, float eps, const void* Wq, const void* Wmeta,
                                 void* out, int64_t ldo, const void* resid, int M, int N, int K, int epi, const float* ss_in,
                                 int n_ss_in, float* ss_out, const void* l2_prefetch, int64_t l2_prefetch_bytes,
                                 cudaStream_t st) {
    P3_CHECK_ARG(Wq && Wmeta, "gemm_skinny_w4: codes and meta are required");
    return skinny_impl(X, ldx, norm_w, eps, nullptr, Wq, Wmeta, out, ldo, resid, M, N, K, epi, ss_in, n_ss_in, ss_out,
                       l2_prefetch, l2_prefetch_bytes, st);
}

// qkv_proj + SuRoPE + paged KV write in one launch (decode, B*L <= 16 tokens): phi.py:442-453.
static int skinny_qkv_rope_impl(const void* X, int64_t ldx, const void* norm_w, float eps, const void* Wqkv, const void* Wq,
                                const void* Wmeta, void* qkv, const float* ss_in, int n_ss_in, const float* cosT, const float* sinT,
                                int64_t tab_bstride, int B, int L, int n_heads, int n_kv, int hd, int K, int past,
                                const int32_t* past_dev, int row_div, void* pool, const int32_t* block_table,
                                int bt_stride, int write_cache, const void* l2_prefetch, int64_t l2_prefetch_bytes,
                                cudaStream_t st, int rs_epi = 0) {
    const int M = B * L;
    P3_CHECK_ARG(!rs_epi || (ss_in && !norm_w), "gemm_skinny_qkv_rope: rs_epi needs ss_in and no norm_w");
    P3_CHECK_ARG(!Wq || (K % 128 == 0 && Wmeta), "gemm_skinny_qkv_rope_w4: K must be a multiple of 128 and meta must be given");
    P3_CHECK_ARG(M >= 1 && M <= 16, "gemm_skinny_qkv_rope: B*L must be in [1,16] (got %d)", M);
    P3_CHECK_ARG(K % 64 == 0 && ldx % 8 == 0, "gemm_skinny_qkv_rope: K %% 64 and ldx %% 8 required");
    P3_CHECK_ARG(hd % 32 == 0 && (n_kv * hd) % 32 == 0, "gemm_skinny_qkv_rope: head_dim must be a multiple of 32");
    P3_CHECK_ARG(!write_cache || (pool && block_table), "gemm_skinny_qkv_rope: cache write needs pool and block table");
    SkParams p{};
    p.X = (const bf16*)X; p.ldx = ldx; p.norm_w = (const bf16*)norm_w; p.eps = eps; p.W = (const bf16*)Wqkv; p.out = qkv;
    p.Wq = (const uint8_t*)Wq; p.Wmeta = (const bf16*)Wmeta;
    p.ldo = (int64_t)(n_heads + 2 * n_kv) * hd; p.M = M; p.N = (n_heads + 2 * n_kv) * hd; p.K = K; p.epi = P3_EPI_ROPE_QKV;
    p.ss_in = ss_in; p.n_ss_in = n_ss_in;
    p.cosT = cosT; p.sinT = sinT; p.tab_bstride = tab_bstride; p.L = L; p.n_heads = n_heads; p.n_kv = n_kv; p.hd = hd;
    p.past = past; p.past_dev = past_dev; p.row_div = row_div; p.write_cache = write_cache;
    p.pool = (bf16*)pool; p.block_table = block_table; p.bt_stride = bt_stride;
    p.l2_pf = (const uint8_t*)l2_prefetch; p.l2_pf_bytes = l2_prefetch_bytes;
    p.rs_epi = rs_epi;
    if (sk_pair_mt1() && M <= 8) return launch_skinny<1, 1>(p, (unsigned)((n_heads + n_kv) * (hd / 16) + n_kv * hd / 16), st);
    unsigned grid = (unsigned)((n_heads + n_kv) * (hd / 32) + n_kv * hd / 32);
    return M <= 8 ? launch_skinny<1, 2>(p, grid, st) : launch_skinny<2, 2>(p, grid, st);
}

extern "C" int p3_gemm_skinny_qkv_rope(const void* X, int64_t ldx, const void* norm_w, float eps, const void* Wqkv,
                                       void* qkv, const float* ss_in, int n_ss_in, const float* cosT, const float* sinT,
                                       int64_t tab_bstride, int B, int L, int n_heads, int n_kv, int hd, int K, int past,
                                       const int32_t* past_dev, int row_div, void* pool, const int32_t* block_table,
                                       int bt_stride, int write_cache, const void* l2_prefetch, int64_t l2_prefetch_bytes,
                                       cudaStream_t st) {
    return skinny_qkv_rope_impl(X, ldx, norm_w, eps, Wqkv, nullptr, nullptr, qkv, ss_in, n_ss_in, cosT, sinT, tab_bstride, B, L,
                                n_heads, n_kv, hd, K, past, past_dev, row_div, pool, block_table, bt_stride, write_cache,
                                l2_prefetch, l2_prefetch_bytes, st);
}

extern "C" int p3_gemm_skinny_qkv_rope_w4(const void* X, int64_t ldx, const void* norm_w, float eps, const void* Wq,
                                          const void* Wmeta, void* qkv, const float* ss_in, int n_ss_in, const float* cosT,
                                          const float* sinT, int64_t tab_bstride, int B, int L, int n_heads, int n_kv, int hd,
                                          int K, int past, const int32_t* past_dev, int row_div, void* pool,
                                          const int32_t* block_table, int bt_stride, int write_cache, const void* l2_prefetch,
                                          int64_t l2_prefetch_bytes, cudaStream_t st) {
    P3_CHECK_ARG(Wq && Wmeta, "gemm_skinny_qkv_rope_w4: codes and meta are required");
    return skinny_qkv_rope_impl(X, ldx, norm_w, eps, nullptr, Wq, Wmeta, qkv, ss_in, n_ss_in, cosT, sinT, tab_bstride, B, L,
                                n_heads, n_kv, hd, K, past, past_dev, row_div, pool, block_table, bt_stride, write_cache,
                                l2_prefetch, l2_prefetch_bytes, st);
}

#include <cstddef>
static_assert(sizeof(p3_skinny_args) == 272 && offsetof(p3_skinny_args, past_dev) == 240 && offsetof(p3_skinny_args, packed) == 264, "p3_skinny_args layout is mirrored by _lib.SkinnyArgs");
// Struct-argument entry for both forms (plain / qkv+rope, bf16 / 4-bit) with the producer-consumer norm split.
extern "C" int p3_gemm_skinny_x(const p3_skinny_args* a, cudaStream_t st) {
    P3_CHECK_ARG(a && (a->op == 0 || a->op == 1), "gemm_skinny_x: op must be 0 (linear) or 1 (qkv + rope)");
    P3_CHECK_ARG((a->W != nullptr) != (a->Wq != nullptr), "gemm_skinny_x: exactly one of W (bf16) and Wq/Wmeta (4-bit) must be given");
    P3_CHECK_ARG(!a->packed || a->op == 0, "gemm_skinny_x: the packed layout applies to op 0 only");
    if (a->op == 1)
        return skinny_qkv_rope_impl(a->X, a->ldx, a->norm_w, a->eps, a->W, a->Wq, a->Wmeta, a->out, a->ss_in, a->n_ss_in, a->cosT, a->sinT,
                                    a->tab_bstride, a->B, a->L, a->n_heads, a->n_kv, a->hd, a->K, a->past, a->past_dev, a->row_div, a->pool,
                                    a->block_table, a->bt_stride, a->write_cache, a->l2_prefetch, a->l2_prefetch_bytes, st, a->rs_epi);
    return skinny_impl(a->X, a->ldx, a->norm_w, a->eps, a->W, a->Wq, a->Wmeta, a->out, a->ldo, a->resid, a->M, a->N, a->K, a->epi, a->ss_in,
                       a->n_ss_in, a->ss_out, a->l2_prefetch, a->l2_prefetch_bytes, st, a->xg_gain, a->xg_out, a->ldxg, a->rs_epi, a->packed);
}
