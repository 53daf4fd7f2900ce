// Skinny (decode-time) GEMM: Y[M,N] = X[M,K] . W[N,K]^T for M <= 16 tokens.
// Replaces nn.Linear qkv_proj / o_proj / gate_up_proj / down_proj / lm_head at decode
// (phi.py:437-438,465-466,604) where the op is a pure weight stream: HBM-bound, so the design
// goal is 16-byte loads of W with many bytes in flight, not tensor-core occupancy.
//
// Mapping: W rows are the MMA "M" dimension (m16n8k16, bf16 -> fp32), the <=16 tokens are "N".
// Each lane loads 16 contiguous bytes of a W row straight from HBM into the A-fragment
// registers; the k-index permutation this implies is applied identically to the X operand, so
// no shared-memory transpose of W is needed. A CTA owns 16*MT W rows and splits K over its
// 8 warps; partial sums are reduced through shared memory.
//
// Fusions: RMSNorm prologue (phi.py:478-479: x*rsqrt(mean(x^2)+eps)*w, rounded to bf16),
// residual epilogue (phi.py:483,485), SwiGLU epilogue (phi.py:470-471), fp32 logits out.
#include "common.cuh"
#include "../../include/phi3_b200.h"

#define SK_WARPS 8
#define SK_THREADS (SK_WARPS * 32)

struct SkParams {
    const bf16* X; int64_t ldx;
    const bf16* norm_w; float eps;
    const bf16* W;
    void* out; int64_t ldo;
    const bf16* resid;
    int M, N, K, epi;
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }

template <int NT, int MT>
__global__ void __launch_bounds__(SK_THREADS, 2) gemm_skinny_kernel(SkParams p) {
    __shared__ float s_rs[16];
    __shared__ float s_red[SK_WARPS][MT][8 * NT][16 + 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int K = p.K, n_chunks = K / 64;

    // W row pointers for this lane (rows g and g+8 of each 16-row tile)
    const bf16* wrow[MT][2];
    int out_col0;
    if (p.epi == P3_EPI_SWIGLU) {
        // interleaved gate/up layout: [128 gate rows | 128 up rows] per 256-row block
        int o0 = blockIdx.x * 16;
        int gate0 = (o0 / 128) * 256 + (o0 % 128);
        out_col0 = o0;
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            wrow[mt][0] = p.W + (size_t)(gate0 + mt * 128 + g) * K + t * 8;
            wrow[mt][1] = p.W + (size_t)(gate0 + mt * 128 + g + 8) * K + t * 8;
        }
    } else {
        int n0 = blockIdx.x * 16 * MT;
        out_col0 = n0;
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            int r0 = min(n0 + mt * 16 + g, p.N - 1), r1 = min(n0 + mt * 16 + g + 8, p.N - 1);
            wrow[mt][0] = p.W + (size_t)r0 * K + t * 8;
            wrow[mt][1] = p.W + (size_t)r1 * K + t * 8;
        }
    }

    // prefetch the first W chunk before the norm prologue so HBM latency overlaps it
    uint4 wreg[MT][2][2];
    int ci = warp;
    if (ci < n_chunks) {
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                wreg[mt][h][0] = ldg_nc_v4(wrow[mt][h] + ci * 64);
                wreg[mt][h][1] = ldg_nc_v4(wrow[mt][h] + ci * 64 + 32);
            }
    }

    // ---- RMSNorm prologue: rs[m] for every token (each CTA recomputes; X is L2 resident)
    if (p.norm_w) {
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

    float acc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[mt][nt][j] = 0.f;

    float rs[NT];
    const bf16* xrow[NT];
    bool xok[NT];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        int m = nt * 8 + g;
        xok[nt] = m < p.M;
        xrow[nt] = p.X + (size_t)(xok[nt] ? m : 0) * p.ldx + t * 8;
        rs[nt] = (p.norm_w && xok[nt]) ? s_rs[m] : 1.f;
    }

    for (; ci < n_chunks; ci += SK_WARPS) {
        uint4 wcur[MT][2][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int h = 0; h < 2; h++) { wcur[mt][h][0] = wreg[mt][h][0]; wcur[mt][h][1] = wreg[mt][h][1]; }
        int cn = ci + SK_WARPS;
        if (cn < n_chunks) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    wreg[mt][h][0] = ldg_nc_v4(wrow[mt][h] + cn * 64);
                    wreg[mt][h][1] = ldg_nc_v4(wrow[mt][h] + cn * 64 + 32);
                }
        }
        // X fragments (same k permutation as W): 2 x 16B per token row per chunk
        uint4 xf[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            if (xok[nt]) {
                xf[nt][0] = *reinterpret_cast<const uint4*>(xrow[nt] + ci * 64);
                xf[nt][1] = *reinterpret_cast<const uint4*>(xrow[nt] + ci * 64 + 32);
            } else {
                xf[nt][0] = make_uint4(0, 0, 0, 0); xf[nt][1] = make_uint4(0, 0, 0, 0);
            }
        }
        if (p.norm_w) {
            uint4 nw[2];
            nw[0] = __ldg(reinterpret_cast<const uint4*>(p.norm_w + ci * 64 + t * 8));
            nw[1] = __ldg(reinterpret_cast<const uint4*>(p.norm_w + ci * 64 + 32 + t * 8));
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t* ux = reinterpret_cast<uint32_t*>(&xf[nt][h]);
                    const uint32_t* uw = reinterpret_cast<const uint32_t*>(&nw[h]);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float2 f = unpack_bf16(ux[j]), w = unpack_bf16(uw[j]);
                        ux[j] = pack_bf16(f.x * rs[nt] * w.x, f.y * rs[nt] * w.y);
                    }
                }
        }
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const uint32_t* w0 = reinterpret_cast<const uint32_t*>(&wcur[mt][0][h]);
                    const uint32_t* w1 = reinterpret_cast<const uint32_t*>(&wcur[mt][1][h]);
                    uint32_t a[4] = {w0[2 * s], w1[2 * s], w0[2 * s + 1], w1[2 * s + 1]};
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        const uint32_t* ux = reinterpret_cast<const uint32_t*>(&xf[nt][h]);
                        mma_bf16_16816(acc[mt][nt], a, ux[2 * s], ux[2 * s + 1]);
                    }
                }
    }

    // ---- cross-warp reduction through shared memory
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

    if (p.epi == P3_EPI_SWIGLU) {
        for (int o = tid; o < 8 * NT * 16; o += SK_THREADS) {
            int r = o & 15, tok = o >> 4;
            if (tok >= p.M) continue;
            float gsum = 0.f, usum = 0.f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; w++) { gsum += s_red[w][0][tok][r]; usum += s_red[w][MT - 1][tok][r]; }
            float gb = bf16_round(gsum), ub = bf16_round(usum);
            float a = bf16_round(silu_f(gb));
            reinterpret_cast<bf16*>(p.out)[(size_t)tok * p.ldo + out_col0 + r] = __float2bfloat16_rn(a * ub);
        }
    } else {
        for (int o = tid; o < MT * 8 * NT * 16; o += SK_THREADS) {
            int r = o & 15, mt = (o >> 4) % MT, tok = o / (16 * MT);
            int n = out_col0 + mt * 16 + r;
            if (tok >= p.M || n >= p.N) continue;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; w++) s += s_red[w][mt][tok][r];
            size_t off = (size_t)tok * p.ldo + n;
            if (p.epi == P3_EPI_F32) {
                reinterpret_cast<float*>(p.out)[off] = s;
            } else if (p.epi == P3_EPI_RESIDUAL) {
                float rv = __bfloat162float(p.resid[off]);
                reinterpret_cast<bf16*>(p.out)[off] = __float2bfloat16_rn(rv + bf16_round(s));
            } else {
                reinterpret_cast<bf16*>(p.out)[off] = __float2bfloat16_rn(s);
            }
        }
    }
}

extern "C" int p3_gemm_skinny(const void* X, int64_t ldx, const void* norm_w, float eps, const void* W, void* out,
                              int64_t ldo, const void* resid, int M, int N, int K, int epi, cudaStream_t st) {
    P3_CHECK_ARG(M >= 1 && M <= 16, "gemm_skinny: M must be in [1,16] (got %d)", M);
    P3_CHECK_ARG(K % 64 == 0, "gemm_skinny: K must be a multiple of 64 (got %d)", K);
    P3_CHECK_ARG(epi == P3_EPI_NONE || epi == P3_EPI_RESIDUAL || epi == P3_EPI_SWIGLU || epi == P3_EPI_F32,
                 "gemm_skinny: unsupported epilogue %d", epi);
    P3_CHECK_ARG(epi != P3_EPI_RESIDUAL || resid, "gemm_skinny: residual epilogue needs resid");
    P3_CHECK_ARG(ldx % 8 == 0, "gemm_skinny: ldx must be a multiple of 8");
    SkParams p{(const bf16*)X, ldx, (const bf16*)norm_w, eps, (const bf16*)W, out, ldo, (const bf16*)resid, M, N, K, epi};
    if (epi == P3_EPI_SWIGLU) {
        P3_CHECK_ARG(N % 256 == 0, "gemm_skinny: SwiGLU needs N (gate+up rows) to be a multiple of 256");
        unsigned grid = (unsigned)(N / 2 / 16);
        if (M <= 8) gemm_skinny_kernel<1, 2><<<grid, SK_THREADS, 0, st>>>(p);
        else gemm_skinny_kernel<2, 2><<<grid, SK_THREADS, 0, st>>>(p);
    } else if (N >= 148 * 32 * 2) {
        unsigned grid = (unsigned)((N + 31) / 32);
        if (M <= 8) gemm_skinny_kernel<1, 2><<<grid, SK_THREADS, 0, st>>>(p);
        else gemm_skinny_kernel<2, 2><<<grid, SK_THREADS, 0, st>>>(p);
    } else {
        unsigned grid = (unsigned)((N + 15) / 16);
        if (M <= 8) gemm_skinny_kernel<1, 1><<<grid, SK_THREADS, 0, st>>>(p);
        else gemm_skinny_kernel<2, 1><<<grid, SK_THREADS, 0, st>>>(p);
    }
    P3_CHECK_LAUNCH("gemm_skinny");
    return 0;
}
