// HBM-bound row kernels: embedding gather, RMSNorm, LayerNorm, SuRoPE + paged-KV write,
// argmax / log-prob / top-k row statistics. Each cites the reference op it replaces.
#include "common.cuh"
#include "../../include/phi3_b200.h"

// ------------------------------------------------------------------------------------------
// embed_gather: nn.Embedding (phi.py:568,577). Negative placeholder ids (phi.py:270) and
// out-of-range ids are clamped to row 0; those rows are overwritten by the image scatter.
// ------------------------------------------------------------------------------------------
__global__ void embed_gather_kernel(const bf16* __restrict__ table, const int32_t* __restrict__ ids,
                                    bf16* __restrict__ out, int64_t T, int H, int vocab, float* __restrict__ ss_out,
                                    const bf16* __restrict__ xg_gain, bf16* __restrict__ xg_out) {
    __shared__ float s_part[4];
    pdl_trigger();
    pdl_wait();
    int64_t t = blockIdx.x;
    int id = ids[t];
    if (id < 0 || id >= vocab) id = 0;
    const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)id * H);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)t * H);
    float ss = 0.f;
    for (int i = threadIdx.x; i < H / 8; i += blockDim.x) {
        uint4 v = __ldg(src + i);
        dst[i] = v;
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
        for (int j = 0; j < 4; j++) { float2 f = unpack_bf16(u[j]); ss += f.x * f.x + f.y * f.y; }
        if (xg_out) {                                        // gain-scaled copy for a consumer that applies rstd in its epilogue
            const uint4 gv = __ldg(reinterpret_cast<const uint4*>(xg_gain) + i);
            const uint32_t* gu = reinterpret_cast<const uint32_t*>(&gv);
            uint4 ov; uint32_t* ou = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
            for (int j = 0; j < 4; j++) { float2 f = unpack_bf16(u[j]), w = unpack_bf16(gu[j]); ou[j] = pack_bf16(f.x * w.x, f.y * w.y); }
            reinterpret_cast<uint4*>(xg_out + (size_t)t * H)[i] = ov;
        }
    }
    if (ss_out) {                                            // sum of squares of the row, for the fused RMSNorm
        ss = warp_sum(ss);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
        __syncthreads();
        if (threadIdx.x == 0) ss_out[t] = (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]);
    }
}

extern "C" int p3_embed_gather(const void* table, const int32_t* ids, void* out, int64_t T, int H, int vocab,
                               float* ss_out, cudaStream_t st) {
    P3_CHECK_ARG(H % 8 == 0, "embed_gather: H must be a multiple of 8");
    if (T == 0) return 0;
    p3_launch_pdl(embed_gather_kernel, dim3((unsigned)T), dim3(128), 0, st, (const bf16*)table, ids, (bf16*)out, T, H, vocab, ss_out,
                  (const bf16*)nullptr, (bf16*)nullptr);
    P3_CHECK_LAUNCH("embed_gather");
    return 0;
}
extern "C" int p3_embed_gather_xg(const void* table, const int32_t* ids, void* out, int64_t T, int H, int vocab, float* ss_out,
                                  const void* xg_gain, void* xg_out, cudaStream_t st) {
    P3_CHECK_ARG(H % 8 == 0 && xg_gain && xg_out, "embed_gather_xg: H must be a multiple of 8; gain and output are required");
    if (T == 0) return 0;
    p3_launch_pdl(embed_gather_kernel, dim3((unsigned)T), dim3(128), 0, st, (const bf16*)table, ids, (bf16*)out, T, H, vocab, ss_out,
                  (const bf16*)xg_gain, (bf16*)xg_out);
    P3_CHECK_LAUNCH("embed_gather_xg");
    return 0;
}

// ------------------------------------------------------------------------------------------
// rope_table: SuRoPE cos/sin table (phi.py:487-507) built on the device. f = fp32(pos) * inv_freq (one fp32 multiply, as the
// reference), full-range cosf / sinf, times the LongRoPE scale. Positions: 0..L_all-1, or per row cat[pids, pids[-1]+1+arange]
// (phi.py:493-497; pad slots carry pid 1). A 128K table is 50 MB: it never crosses PCIe.
// ------------------------------------------------------------------------------------------
__global__ void rope_table_kernel(const int32_t* __restrict__ pids, int64_t pid_stride, int Lp, const float* __restrict__ inv_freq,
                                  float* __restrict__ cosT, float* __restrict__ sinT, int64_t total, int L_all, int half, float sf) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int d = (int)(idx % half);
    const int i = (int)((idx / half) % L_all);
    const int64_t b = idx / ((int64_t)half * L_all);
    int pos = i;
    if (pids) pos = i < Lp ? pids[b * pid_stride + i] : pids[b * pid_stride + Lp - 1] + 1 + (i - Lp);
    const float f = __fmul_rn((float)pos, inv_freq[d]);
    cosT[idx] = __fmul_rn(cosf(f), sf);
    sinT[idx] = __fmul_rn(sinf(f), sf);
}

extern "C" int p3_rope_table(const int32_t* pids, int64_t pid_stride, int Lp, const float* inv_freq, float* cosT, float* sinT, int Bt,
                             int L_all, int half, float scale, cudaStream_t st) {
    P3_CHECK_ARG(Bt >= 1 && L_all >= 1 && half >= 1 && (!pids || (Lp >= 1 && Lp <= L_all)), "rope_table: bad sizes");
    const int64_t total = (int64_t)Bt * L_all * half;
    rope_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pids, pid_stride, Lp, inv_freq, cosT, sinT, total, L_all, half, scale);
    P3_CHECK_LAUNCH("rope_table");
    return 0;
}

// ------------------------------------------------------------------------------------------
// row_sumsq: sum of squares of every row (the statistics half of nn.RMSNorm, phi.py:478): one warp per row.
// ------------------------------------------------------------------------------------------
__global__ void row_sumsq_kernel(const bf16* __restrict__ x, int64_t ldx, float* __restrict__ out, int64_t T, int H) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    pdl_trigger();
    pdl_wait();
    if (row >= T) return;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * ldx);
    float ss = 0.f;
    for (int c = lane; c < H / 8; c += 32) {
        const uint4 v = xr[c];
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
        for (int j = 0; j < 4; j++) { const float2 f = unpack_bf16(u[j]); ss += f.x * f.x + f.y * f.y; }
    }
    ss = warp_sum(ss);
    if (lane == 0) out[row] = ss;
}

extern "C" int p3_row_sumsq(const void* x, int64_t ldx, float* out, int64_t T, int H, cudaStream_t st) {
    P3_CHECK_ARG(H % 8 == 0 && ldx % 8 == 0, "row_sumsq: H and ldx must be multiples of 8");
    if (T == 0) return 0;
    p3_launch_pdl(row_sumsq_kernel, dim3((unsigned)((T + 7) / 8)), dim3(256), 0, st, (const bf16*)x, ldx, out, T, H);
    P3_CHECK_LAUNCH("row_sumsq");
    return 0;
}

// ------------------------------------------------------------------------------------------
// rmsnorm: nn.RMSNorm -> mx.fast.rms_norm (phi.py:478-479,571): fp32 accumulate, one rounding.
// One warp per row, row held in registers (H <= 8192).
// ------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void rmsnorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                               int64_t T, int H, float eps) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    pdl_trigger();
    pdl_wait();
    if (row >= T) return;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * H);
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * H);
    int nv = H / 8;
    uint4 v[MAXV];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
        int c = lane + i * 32;
        if (c < nv) {
            v[i] = xr[c];
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
            for (int j = 0; j < 4; j++) { float2 f = unpack_bf16(u[j]); ss += f.x * f.x + f.y * f.y; }
        }
    }
    ss = warp_sum(ss);
    float rs = rsqrtf(ss / (float)H + eps);
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
        int c = lane + i * 32;
        if (c < nv) {
            uint4 wv = __ldg(wr + c), o;
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[i]);
            const uint32_t* uw = reinterpret_cast<const uint32_t*>(&wv);
            uint32_t* uo = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float2 f = unpack_bf16(u[j]), g = unpack_bf16(uw[j]);
                uo[j] = pack_bf16(f.x * rs * g.x, f.y * rs * g.y);
            }
            yr[c] = o;
        }
    }
}

// Decode-time variant (T <= 16 rows): one CTA per row, programmatic dependent launch on both sides, so that the
// skinny GEMMs that follow stream weights only (normalising X inside every 32-row GEMM CTA costs more
// instructions than its weight tile: tools/microbench.py, qkv 14.9 -> 11.1 us without the fused norm).
__global__ void __launch_bounds__(256) rmsnorm_rows_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                          bf16* __restrict__ y, int H, float eps) {
    __shared__ float s_part[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nv = H / 8;
    pdl_trigger();
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    uint4 wv[4];
#pragma unroll
    for (int i = 0; i < 4; i++) wv[i] = (tid + i * 256 < nv) ? __ldg(wr + tid + i * 256) : make_uint4(0, 0, 0, 0);   // gains: immutable
    pdl_wait();
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)blockIdx.x * H);
    uint4* yr = reinterpret_cast<uint4*>(y + (size_t)blockIdx.x * H);
    uint4 v[4];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int c = tid + i * 256;
        if (c < nv) {
            v[i] = xr[c];
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
            for (int j = 0; j < 4; j++) { float2 f = unpack_bf16(u[j]); ss += f.x * f.x + f.y * f.y; }
        }
    }
    ss = warp_sum(ss);
    if (lane == 0) s_part[warp] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) tot += s_part[i];
    const float rs = rsqrtf(tot / (float)H + eps);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int c = tid + i * 256;
        if (c < nv) {
            uint4 o;
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[i]);
            const uint32_t* uw = reinterpret_cast<const uint32_t*>(&wv[i]);
            uint32_t* uo = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float2 f = unpack_bf16(u[j]), g = unpack_bf16(uw[j]);
                uo[j] = pack_bf16(f.x * rs * g.x, f.y * rs * g.y);
            }
            yr[c] = o;
        }
    }
}

extern "C" int p3_rmsnorm(const void* x, const void* w, void* y, int64_t T, int H, float eps, cudaStream_t st) {
    P3_CHECK_ARG(H % 8 == 0 && H <= 8192, "rmsnorm: H must be a multiple of 8 and <= 8192");
    if (T == 0) return 0;
    if (T <= 16) {
        p3_launch_pdl(rmsnorm_rows_kernel, dim3((unsigned)T), dim3(256), (size_t)0, st, (const bf16*)x, (const bf16*)w, (bf16*)y, H, eps);
        P3_CHECK_LAUNCH("rmsnorm_rows");
        return 0;
    }
    unsigned grid = (unsigned)((T + 3) / 4);
    if (H <= 4096)
        p3_launch_pdl(rmsnorm_kernel<16>, dim3(grid), dim3(128), (size_t)0, st, (const bf16*)x, (const bf16*)w, (bf16*)y, T, H, eps);
    else
        p3_launch_pdl(rmsnorm_kernel<32>, dim3(grid), dim3(128), (size_t)0, st, (const bf16*)x, (const bf16*)w, (bf16*)y, T, H, eps);
    P3_CHECK_LAUNCH("rmsnorm");
    return 0;
}

// ------------------------------------------------------------------------------------------
// layernorm: nn.LayerNorm -> mx.fast.layer_norm (phi.py:165,167,212), eps 1e-5, with bias.
// Input is the fp32 CLIP residual stream (the reference runs CLIP activations in fp32);
// output bf16 (feeds a GEMM) or fp32 (pre_layrnorm, phi.py:218). One warp per row, H <= 1024.
// ------------------------------------------------------------------------------------------
template <bool OUT_F32>
__global__ void layernorm_kernel(const float* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                                 void* __restrict__ y, int64_t T, int H, float eps) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    pdl_trigger();
    pdl_wait();
    if (row >= T) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * H);
    int nv = H / 4;
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int c = lane + i * 32;
        if (c < nv) { v[i] = xr[c]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
    }
    float mean = warp_sum(s) / (float)H;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int c = lane + i * 32;
        if (c < nv) {
            float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
            ss += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
        }
    }
    float rs = rsqrtf(warp_sum(ss) / (float)H + eps);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int c = lane + i * 32;
        if (c < nv) {
            uint2 wv = __ldg(reinterpret_cast<const uint2*>(w) + c), bv = __ldg(reinterpret_cast<const uint2*>(b) + c);
            float2 w0 = unpack_bf16(wv.x), w1 = unpack_bf16(wv.y), b0 = unpack_bf16(bv.x), b1 = unpack_bf16(bv.y);
            float o0 = (v[i].x - mean) * rs * w0.x + b0.x, o1 = (v[i].y - mean) * rs * w0.y + b0.y;
            float o2 = (v[i].z - mean) * rs * w1.x + b1.x, o3 = (v[i].w - mean) * rs * w1.y + b1.y;
            if (OUT_F32) {
                reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (size_t)row * H)[c] = make_float4(o0, o1, o2, o3);
            } else {
                reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + (size_t)row * H)[c] = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
            }
        }
    }
}

extern "C" int p3_layernorm(const float* x, const void* w, const void* b, void* y, int64_t T, int H, float eps,
                            int out_f32, cudaStream_t st) {
    P3_CHECK_ARG(H % 4 == 0 && H <= 1024, "layernorm: H must be a multiple of 4 and <= 1024");
    if (T == 0) return 0;
    unsigned grid = (unsigned)((T + 3) / 4);
    if (out_f32) p3_launch_pdl(layernorm_kernel<true>, dim3(grid), dim3(128), (size_t)0, st, x, (const bf16*)w, (const bf16*)b, y, T, H, eps);
    else p3_launch_pdl(layernorm_kernel<false>, dim3(grid), dim3(128), (size_t)0, st, x, (const bf16*)w, (const bf16*)b, y, T, H, eps);
    P3_CHECK_LAUNCH("layernorm");
    return 0;
}

// ------------------------------------------------------------------------------------------
// rope_kvwrite: _rotate_half on q,k (phi.py:418-423,451-452) with the SuRoPE table
// (phi.py:487-507) + KVCache slice-assign (phi.py:542-548) into the paged pool.
// qkv: [B*L, (n_heads + 2*n_kv)*hd] bf16, roped in place. cos/sin: fp32 [Bt, L_all, hd/2]
// (half table: the reference concatenates [freqs, freqs]); row b uses table row (b / row_div)
// when tab_bstride != 0. Token i of row b sits at absolute position past + i.
// ------------------------------------------------------------------------------------------
__global__ void rope_kvwrite_kernel(bf16* __restrict__ qkv, const float* __restrict__ cosT, const float* __restrict__ sinT,
                                    int64_t tab_bstride, int B, int L, int n_heads, int n_kv, int hd, int past,
                                    int row_div, bf16* __restrict__ pool, const int32_t* __restrict__ block_table,
                                    int bt_stride, int write_cache, const int32_t* __restrict__ past_dev) {
    pdl_trigger();
    pdl_wait();
    if (past_dev) past = *past_dev;
    const int half = hd / 2, cpr = half / 8;                  // 8-elem chunks per half head
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)B * L * n_heads * cpr;
    if (idx >= total) return;
    int c = (int)(idx % cpr);
    int h = (int)((idx / cpr) % n_heads);
    int64_t tok = idx / ((int64_t)cpr * n_heads);
    int b = (int)(tok / L), i = (int)(tok % L);
    int pos = past + i;
    int qkv_dim = (n_heads + 2 * n_kv) * hd;
    const float* cr = cosT + (size_t)(b / row_div) * tab_bstride + (size_t)pos * half + c * 8;
    const float* sr = sinT + (size_t)(b / row_div) * tab_bstride + (size_t)pos * half + c * 8;
    float cs[8], sn[8];
    *reinterpret_cast<float4*>(cs) = *reinterpret_cast<const float4*>(cr);
    *reinterpret_cast<float4*>(cs + 4) = *reinterpret_cast<const float4*>(cr + 4);
    *reinterpret_cast<float4*>(sn) = *reinterpret_cast<const float4*>(sr);
    *reinterpret_cast<float4*>(sn + 4) = *reinterpret_cast<const float4*>(sr + 4);
    bf16* row = qkv + (size_t)tok * qkv_dim;

    auto rope8 = [&](bf16* p, uint4& o1, uint4& o2) {
        uint4 a = *reinterpret_cast<uint4*>(p + c * 8), bb = *reinterpret_cast<uint4*>(p + half + c * 8);
        const uint32_t* ua = reinterpret_cast<const uint32_t*>(&a);
        const uint32_t* ub = reinterpret_cast<const uint32_t*>(&bb);
        uint32_t* u1 = reinterpret_cast<uint32_t*>(&o1);
        uint32_t* u2 = reinterpret_cast<uint32_t*>(&o2);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float2 x1 = unpack_bf16(ua[j]), x2 = unpack_bf16(ub[j]);
            // out[:half] = x1*cos - x2*sin ; out[half:] = x2*cos + x1*sin
            u1[j] = pack_bf16(x1.x * cs[2 * j] - x2.x * sn[2 * j], x1.y * cs[2 * j + 1] - x2.y * sn[2 * j + 1]);
            u2[j] = pack_bf16(x2.x * cs[2 * j] + x1.x * sn[2 * j], x2.y * cs[2 * j + 1] + x1.y * sn[2 * j + 1]);
        }
        *reinterpret_cast<uint4*>(p + c * 8) = o1;
        *reinterpret_cast<uint4*>(p + half + c * 8) = o2;
    };
    uint4 o1, o2;
    rope8(row + h * hd, o1, o2);                              // q head h
    if (h < n_kv) {
        rope8(row + (n_heads + h) * hd, o1, o2);              // k head h
        if (write_cache) {
            int page = block_table[(size_t)(b / row_div) * bt_stride + pos / P3_PAGE];
            int slot = pos % P3_PAGE;
            bf16* kd = pool + (size_t)page * kv_page_elems(n_kv, hd) + ((size_t)h * P3_PAGE + slot) * hd;
            bf16* vd = kd + (size_t)n_kv * P3_PAGE * hd;
            *reinterpret_cast<uint4*>(kd + c * 8) = o1;
            *reinterpret_cast<uint4*>(kd + half + c * 8) = o2;
            const bf16* vs = row + (n_heads + n_kv + h) * hd;
            *reinterpret_cast<uint4*>(vd + c * 8) = *reinterpret_cast<const uint4*>(vs + c * 8);
            *reinterpret_cast<uint4*>(vd + half + c * 8) = *reinterpret_cast<const uint4*>(vs + half + c * 8);
        }
    }
}

extern "C" int p3_rope_kvwrite(void* qkv, const float* cosT, const float* sinT, int64_t tab_bstride, int B, int L,
                               int n_heads, int n_kv, int hd, int past, int row_div, void* pool,
                               const int32_t* block_table, int bt_stride, int write_cache, const int32_t* past_dev,
                               cudaStream_t st) {
    P3_CHECK_ARG(hd % 16 == 0, "rope_kvwrite: head_dim must be a multiple of 16");
    P3_CHECK_ARG(n_kv <= n_heads && row_div >= 1, "rope_kvwrite: bad head counts / row_div");
    int64_t total = (int64_t)B * L * n_heads * (hd / 16);
    if (total == 0) return 0;
    p3_launch_pdl(rope_kvwrite_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st,
                  (bf16*)qkv, cosT, sinT, tab_bstride, B, L, n_heads, n_kv, hd, past, row_div, (bf16*)pool, block_table,
                  bt_stride, write_cache, past_dev);
    P3_CHECK_LAUNCH("rope_kvwrite");
    return 0;
}

// ------------------------------------------------------------------------------------------
// row_stats: mx.argmax (pv:386,392,506,559), nn.log_softmax + gathers (pv:476,541-547,571-573),
// mx.argpartition top-n (pv:507) in ONE pass over each logits row.
// Per row r of logits[R, V] (fp32, row stride ld):
//   argmax[r] (first index on ties), maxv[r], lse[r] = log sum exp
//   topk_ids[r, k], topk_lp[r, k]  (k < n_top, descending; log-probs)
//   gather_lp[r, g] = logits[r, gather_ids[r, g]] - lse[r]
// ------------------------------------------------------------------------------------------
#define RS_THREADS 256
__global__ void row_stats_kernel(const float* __restrict__ logits, int64_t ld, int V, int32_t* __restrict__ argmax_out,
                                 float* __restrict__ max_out, float* __restrict__ lse_out, int n_top,
                                 int32_t* __restrict__ topk_ids, float* __restrict__ topk_lp, int n_gather,
                                 const int32_t* __restrict__ gather_ids, float* __restrict__ gather_lp) {
    __shared__ float s_val[RS_THREADS / 32];
    __shared__ int s_idx[RS_THREADS / 32];
    __shared__ float s_sum[RS_THREADS / 32];
    __shared__ int s_taken[8];
    __shared__ float s_bval;
    __shared__ int s_bidx;
    pdl_trigger();
    pdl_wait();
    int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* row = logits + (size_t)r * ld;
    float lse = 0.f, rowmax = 0.f;
    for (int k = 0; k < (n_top > 0 ? n_top : 1); k++) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = tid; i < V; i += RS_THREADS) {
            float v = row[i];
            bool taken = false;
            for (int j = 0; j < k; j++) taken |= (s_taken[j] == i);
            if (!taken && (v > best || (v == best && i < bi))) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < RS_THREADS / 32; w++)
                if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
            s_bval = best; s_bidx = bi;
            if (k < 8) s_taken[k] = bi;
        }
        __syncthreads();
        best = s_bval; bi = s_bidx;
        if (k == 0) {
            rowmax = best;
            float sum = 0.f;
            for (int i = tid; i < V; i += RS_THREADS) sum += __expf(row[i] - rowmax);
            sum = warp_sum(sum);
            if (lane == 0) s_sum[warp] = sum;
            __syncthreads();
            sum = 0.f;
            for (int w = 0; w < RS_THREADS / 32; w++) sum += s_sum[w];
            lse = rowmax + logf(sum);
            if (tid == 0) {
                if (argmax_out) argmax_out[r] = bi;
                if (max_out) max_out[r] = rowmax;
                if (lse_out) lse_out[r] = lse;
            }
        }
        if (n_top > 0 && tid == 0) {
            topk_ids[(size_t)r * n_top + k] = bi;
            topk_lp[(size_t)r * n_top + k] = best - lse;
        }
        __syncthreads();
    }
    for (int g = tid; g < n_gather; g += RS_THREADS) {
        int id = gather_ids[(size_t)r * n_gather + g];
        gather_lp[(size_t)r * n_gather + g] = (id >= 0 && id < V) ? row[id] - lse : -INFINITY;
    }
}

extern "C" int p3_row_stats(const float* logits, int64_t R, int64_t ld, int V, int32_t* argmax_out, float* max_out,
                            float* lse_out, int n_top, int32_t* topk_ids, float* topk_lp, int n_gather,
                            const int32_t* gather_ids, float* gather_lp, cudaStream_t st) {
    P3_CHECK_ARG(n_top >= 0 && n_top <= 8, "row_stats: n_top must be in [0,8]");
    P3_CHECK_ARG(n_top == 0 || (topk_ids && topk_lp), "row_stats: top-k outputs missing");
    P3_CHECK_ARG(n_gather == 0 || (gather_ids && gather_lp), "row_stats: gather buffers missing");
    if (R == 0) return 0;
    p3_launch_pdl(row_stats_kernel, dim3((unsigned)R), dim3(RS_THREADS), 0, st, logits, ld, V, argmax_out, max_out, lse_out,
                  n_top, topk_ids, topk_lp, n_gather, gather_ids, gather_lp);
    P3_CHECK_LAUNCH("row_stats");
    return 0;
}

// ------------------------------------------------------------------------------------------
// decode_advance: greedy-loop bookkeeping kept on the device so a whole decode step replays as
// one CUDA graph with no host sync (the reference syncs twice per token, pv:393,397).
//   history[b][*step] = tok[b];  eos_seen[b] |= tok[b]==eos;  (*step)++;  (*past)++
// ------------------------------------------------------------------------------------------
__global__ void decode_advance_kernel(const int32_t* __restrict__ tok, int32_t* __restrict__ history, int64_t ld, int B,
                                      int32_t* step, int32_t* past, int32_t* eos_seen) {
    pdl_trigger();
    pdl_wait();
    int s = *step;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        history[(size_t)b * ld + s] = tok[b];
        if (eos_seen && tok[b] == 32007) eos_seen[b] = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) { *step = s + 1; if (past) *past = *past + 1; }
}

extern "C" int p3_decode_advance(const int32_t* tok, int32_t* history, int64_t ld, int B, int32_t* step, int32_t* past,
                                 int32_t* eos_seen, cudaStream_t st) {
    p3_launch_pdl(decode_advance_kernel, dim3(1), dim3(128), 0, st, tok, history, ld, B, step, past, eos_seen);
    P3_CHECK_LAUNCH("decode_advance");
    return 0;
}

// ------------------------------------------------------------------------------------------
// top_p_sample: nucleus sampling (north_star "top-p kernel"; the reference is greedy only, pv:386,392 —
// this is an extension, SURVEY.md H13, validated against a torch restatement of the same rule).
// Per row: p = softmax(logits / temperature); tau = the largest probability value v such that
// sum_{p_i >= v} p_i >= top_p (exact: bisection over the float bit pattern, no sort); the token is
// drawn by inverse CDF in INDEX order over the nucleus {i : p_i >= tau} with the supplied uniform u.
// ------------------------------------------------------------------------------------------
#define TP_THREADS 256
__device__ __forceinline__ float block_sum_tp(float v, float* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < TP_THREADS / 32; w++) t += sh[w];
    return t;
}

__global__ void top_p_sample_kernel(const float* __restrict__ logits, int64_t ld, int V, float top_p, float inv_temp,
                                    const float* __restrict__ u, int32_t* __restrict__ out, float* __restrict__ tau_out,
                                    const int32_t* __restrict__ step_dev, int64_t u_stride) {
    __shared__ float sh[TP_THREADS / 32];
    __shared__ float s_chunk[TP_THREADS];
    __shared__ int s_pick;
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x, tid = threadIdx.x;
    const float* row = logits + (size_t)r * ld;
    if (step_dev) u += (size_t)(*step_dev) * u_stride;       // decode graph: row of the uniform table for this step
    float mx = -INFINITY;
    for (int i = tid; i < V; i += TP_THREADS) mx = fmaxf(mx, row[i] * inv_temp);
    mx = warp_max(mx);
    __syncthreads();
    if ((tid & 31) == 0) sh[tid >> 5] = mx;
    __syncthreads();
    mx = sh[0];
#pragma unroll
    for (int w = 1; w < TP_THREADS / 32; w++) mx = fmaxf(mx, sh[w]);
    float z = 0.f;
    for (int i = tid; i < V; i += TP_THREADS) z += __expf(row[i] * inv_temp - mx);
    z = block_sum_tp(z, sh);
    const float inv_z = 1.f / z;
    // bisection over bit patterns of positive floats in (0, 1]
    uint32_t lo = 0u, hi = __float_as_uint(1.0f);           // f(lo) >= top_p always; find the largest v with f(v) >= top_p
    if (top_p >= 1.f) hi = 0u;                               // whole vocabulary: fp32 mass can round below 1
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo + 1) / 2;
        float thr = __uint_as_float(mid), m = 0.f;
        for (int i = tid; i < V; i += TP_THREADS) { float p = __expf(row[i] * inv_temp - mx) * inv_z; if (p >= thr) m += p; }
        m = block_sum_tp(m, sh);
        if (m >= top_p) lo = mid; else hi = mid - 1;
    }
    const float tau = __uint_as_float(lo);
    // inverse CDF in index order over the nucleus: contiguous chunk per thread
    const int per = (V + TP_THREADS - 1) / TP_THREADS, i0 = tid * per, i1 = min(V, i0 + per);
    float cs = 0.f;
    for (int i = i0; i < i1; i++) { float p = __expf(row[i] * inv_temp - mx) * inv_z; if (p >= tau) cs += p; }
    s_chunk[tid] = cs;
    if (tid == 0) s_pick = -1;
    __syncthreads();
    if (tid == 0) {
        float mass = 0.f;
        int last_ne = 0;                                     // last chunk that holds a nucleus element (trailing chunks can be empty)
        for (int t = 0; t < TP_THREADS; t++) { mass += s_chunk[t]; if (s_chunk[t] > 0.f) last_ne = t; }
        float target = u[r] * mass, acc = 0.f;
        int t = 0;
        for (; t < last_ne; t++) { if (acc + s_chunk[t] > target) break; acc += s_chunk[t]; }   // u -> 1 rounds into the last non-empty chunk
        s_pick = t;
        s_chunk[0] = target - acc;                           // residual target inside the chosen chunk (slot 0 reused)
    }
    __syncthreads();
    if (tid == s_pick) {
        float target = s_chunk[0], acc = 0.f;
        int last = -1, pick = -1;
        for (int i = i0; i < i1; i++) {
            float p = __expf(row[i] * inv_temp - mx) * inv_z;
            if (p >= tau) { last = i; acc += p; if (pick < 0 && acc > target) pick = i; }
        }
        out[r] = pick >= 0 ? pick : last;
        if (tau_out) tau_out[r] = tau;
    }
}

extern "C" int p3_top_p_sample(const float* logits, int64_t R, int64_t ld, int V, float top_p, float temperature,
                               const float* u, int32_t* out, float* tau_out, const int32_t* step_dev, int64_t u_stride,
                               cudaStream_t st) {
    P3_CHECK_ARG(top_p > 0.f && top_p <= 1.f && temperature > 0.f, "top_p_sample: need 0 < top_p <= 1 and temperature > 0");
    if (R == 0) return 0;
    p3_launch_pdl(top_p_sample_kernel, dim3((unsigned)R), dim3(TP_THREADS), 0, st, logits, ld, V, top_p, 1.f / temperature, u,
                  out, tau_out, step_dev, u_stride);
    P3_CHECK_LAUNCH("top_p_sample");
    return 0;
}
