// Vision front end: HD transform (PIL-exact bilinear resize, white pad, normalise, crop tiling,
// the reference's 2-tap "interpolate_336" global crop), patch im2col, CLIP embedding assembly and
// the glb_GN / sub_GN token assembly. Integer / index work is bit-exact against the reference's
// numpy/PIL code (phi.py:283-372, 393-416); all of it is HBM-bound byte shuffling.
#include "common.cuh"
#include "../../include/phi3_b200.h"

#define PRECISION_BITS 22   // Pillow Resample.c: 32 - 8 - 2

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: tmp[y][xx][c] = clip8(0.5 + sum_x src[y][xmin+x][c] * kk[xx][x])
__global__ void hd_resize_h_kernel(const uint8_t* __restrict__ src, int64_t sy, int64_t sx, int in_h, uint8_t* __restrict__ tmp,
                                   int out_w, const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)in_h * out_w) return;
    int xx = (int)(idx % out_w), y = (int)(idx / out_w);
    int xmin = bounds[2 * xx], xn = bounds[2 * xx + 1];
    const int32_t* k = kk + (size_t)xx * ksize;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    const uint8_t* row = src + (size_t)y * sy;
    for (int x = 0; x < xn; x++) {
        const uint8_t* px = row + (size_t)(xmin + x) * sx;
        int w = k[x];
        s0 += px[0] * w; s1 += px[1] * w; s2 += px[2] * w;
    }
    uint8_t* o = tmp + idx * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

extern "C" int p3_hd_resize_h(const uint8_t* src, int64_t sy, int64_t sx, int in_w, int in_h, uint8_t* tmp, int out_w,
                              const int32_t* bounds, const int32_t* kk, int ksize, cudaStream_t st) {
    (void)in_w;
    int64_t n = (int64_t)in_h * out_w;
    if (n == 0) return 0;
    hd_resize_h_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, sy, sx, in_h, tmp, out_w, bounds, kk, ksize);
    P3_CHECK_LAUNCH("hd_resize_h");
    return 0;
}

// vertical pass + white padding (phi.py:300-306) + optional un-transpose (phi.py:308)
__global__ void hd_resize_v_pad_kernel(const uint8_t* __restrict__ tmp, int tmp_w, int out_h, const int32_t* __restrict__ bounds,
                                       const int32_t* __restrict__ kk, int ksize, int pad_top, int padded_h, int transposed,
                                       uint8_t* __restrict__ out) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)padded_h * tmp_w) return;
    int x = (int)(idx % tmp_w), yp = (int)(idx / tmp_w);
    int yy = yp - pad_top;
    uint8_t r = 255, g = 255, b = 255;
    if (yy >= 0 && yy < out_h) {
        int ymin = bounds[2 * yy], yn = bounds[2 * yy + 1];
        const int32_t* k = kk + (size_t)yy * ksize;
        int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int y = 0; y < yn; y++) {
            const uint8_t* px = tmp + ((size_t)(ymin + y) * tmp_w + x) * 3;
            int w = k[y];
            s0 += px[0] * w; s1 += px[1] * w; s2 += px[2] * w;
        }
        r = clip8(s0); g = clip8(s1); b = clip8(s2);
    }
    // final image: [padded_h, tmp_w] or, transposed back, [tmp_w, padded_h]
    size_t o = transposed ? ((size_t)x * padded_h + yp) : ((size_t)yp * tmp_w + x);
    out[o * 3] = r; out[o * 3 + 1] = g; out[o * 3 + 2] = b;
}

extern "C" int p3_hd_resize_v_pad(const uint8_t* tmp, int tmp_w, int tmp_h, int out_h, const int32_t* bounds,
                                  const int32_t* kk, int ksize, int pad_top, int padded_h, int transposed,
                                  uint8_t* out_hwc, cudaStream_t st) {
    (void)tmp_h;
    P3_CHECK_ARG(padded_h >= out_h + pad_top, "hd_resize_v_pad: padded_h too small");
    int64_t n = (int64_t)padded_h * tmp_w;
    if (n == 0) return 0;
    hd_resize_v_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tmp, tmp_w, out_h, bounds, kk, ksize, pad_top,
                                                                         padded_h, transposed, out_hwc);
    P3_CHECK_LAUNCH("hd_resize_v_pad");
    return 0;
}

// pixel_values[crop][c][y][x]; crop 0 = global (phi.py:331-372), crops 1.. row-major grid (phi.py:322-326)
__global__ void hd_tile_crops_kernel(const uint8_t* __restrict__ img, int H, int W, const double* __restrict__ lut,
                                     float* __restrict__ pv, const int32_t* __restrict__ h_idx, const float* __restrict__ h_wgt,
                                     const int32_t* __restrict__ w_idx, const float* __restrict__ w_wgt) {
    const int wc = W / 336, hc = H / 336;
    int64_t total = (int64_t)(hc * wc + 1) * 3 * 336 * 336;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = (int)(idx % 336), y = (int)((idx / 336) % 336), c = (int)((idx / (336 * 336)) % 3);
    int crop = (int)(idx / (3 * 336 * 336));
    if (crop > 0) {
        int ci = (crop - 1) / wc, cj = (crop - 1) % wc;
        uint8_t v = img[((size_t)(ci * 336 + y) * W + cj * 336 + x) * 3 + c];
        pv[idx] = (float)lut[v * 3 + c];
        return;
    }
    // out = sum_{a,b} (float)(wh[a]*ww[b]) * in[hi[a]][wi[b]]  — fp32 weight product, float64 sum (numpy semantics)
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 2; a++) {
        int yi = h_idx[2 * y + a];
        float wa = h_wgt[2 * y + a];
#pragma unroll
        for (int b = 0; b < 2; b++) {
            int xi = w_idx[2 * x + b];
            float wp = __fmul_rn(wa, w_wgt[2 * x + b]);
            acc += (double)wp * lut[img[((size_t)yi * W + xi) * 3 + c] * 3 + c];
        }
    }
    pv[idx] = (float)acc;
}

extern "C" int p3_hd_tile_crops(const uint8_t* img_hwc, int H, int W, const double* lut, float* pixel_values,
                                const int32_t* h_idx, const float* h_wgt, const int32_t* w_idx, const float* w_wgt,
                                cudaStream_t st) {
    P3_CHECK_ARG(H % 336 == 0 && W % 336 == 0 && H > 0 && W > 0, "hd_tile_crops: H and W must be multiples of 336");
    int64_t total = (int64_t)((H / 336) * (W / 336) + 1) * 3 * 336 * 336;
    hd_tile_crops_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(img_hwc, H, W, lut, pixel_values, h_idx, h_wgt,
                                                                           w_idx, w_wgt);
    P3_CHECK_LAUNCH("hd_tile_crops");
    return 0;
}

// im2col for the 14x14 stride-14 conv: A[(n*576 + py*24+px)][(ky*14+kx)*3 + c]
__global__ void patch_im2col_kernel(const float* __restrict__ pv, bf16* __restrict__ A, int N, int Kpad) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)N * 576 * Kpad;
    if (idx >= total) return;
    int k = (int)(idx % Kpad);
    int64_t row = idx / Kpad;
    float v = 0.f;
    if (k < 588) {
        int c = k % 3, kx = (k / 3) % 14, ky = k / 42;
        int p = (int)(row % 576), n = (int)(row / 576);
        int py = p / 24, px = p % 24;
        v = pv[(((size_t)n * 3 + c) * 336 + py * 14 + ky) * 336 + px * 14 + kx];
    }
    A[idx] = __float2bfloat16_rn(v);
}

extern "C" int p3_patch_im2col(const float* pixel_values, void* A, int N, int Kpad, cudaStream_t st) {
    P3_CHECK_ARG(Kpad >= 588 && Kpad % 8 == 0, "patch_im2col: Kpad must be >= 588 and a multiple of 8");
    int64_t total = (int64_t)N * 576 * Kpad;
    if (total == 0) return 0;
    patch_im2col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pixel_values, (bf16*)A, N, Kpad);
    P3_CHECK_LAUNCH("patch_im2col");
    return 0;
}

// embeddings = cat[cls, patches] + pos (phi.py:202-205), fp32
__global__ void clip_embed_kernel(const float* __restrict__ patches, const bf16* __restrict__ cls, const bf16* __restrict__ pos,
                                  float* __restrict__ out, int N, int D) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)N * 577 * D;
    if (idx >= total) return;
    int d = (int)(idx % D);
    int tok = (int)((idx / D) % 577);
    int n = (int)(idx / ((int64_t)D * 577));
    float v = tok == 0 ? __bfloat162float(cls[d]) : patches[((size_t)n * 576 + tok - 1) * D + d];
    out[idx] = v + __bfloat162float(pos[(size_t)tok * D + d]);
}

extern "C" int p3_clip_embed(const float* patches, const void* cls, const void* pos, float* out, int N, int D,
                             cudaStream_t st) {
    int64_t total = (int64_t)N * 577 * D;
    if (total == 0) return 0;
    clip_embed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(patches, (const bf16*)cls, (const bf16*)pos, out, N, D);
    P3_CHECK_LAUNCH("clip_embed");
    return 0;
}

// Token assembly for one image (phi.py:403-407, closed form in SURVEY.md A.3):
//   [ sub tokens (plain-reshape order) with a sub_GN after every 12*wc tokens | glb_GN | global 12 rows of 12 + sub_GN ]
// merged token (crop, r, c) channel block q=(dr*2+dc) <- feature (2r+dr, 2c+dc) of the crop (CLS dropped).
__global__ void gn_assemble_kernel(const float* __restrict__ feats, const bf16* __restrict__ sub_GN, const bf16* __restrict__ glb_GN,
                                   bf16* __restrict__ out, int hc, int wc, int C) {
    const int n_sub = hc * wc * 144;
    const int sub_len = n_sub + 12 * hc;                       // with separators
    const int total_tok = sub_len + 1 + 156;
    int tok = blockIdx.x;
    if (tok >= total_tok) return;
    bf16* o = out + (size_t)tok * 4 * C;
    const bf16* gn = nullptr;
    int crop = 0, r = 0, c = 0;
    if (tok < sub_len) {
        int rowlen = 12 * wc + 1;
        int row = tok / rowlen, col = tok % rowlen;
        if (col == rowlen - 1) gn = sub_GN;
        else { int t = row * (12 * wc) + col; crop = 1 + t / 144; r = (t % 144) / 12; c = t % 12; }
    } else if (tok == sub_len) {
        gn = glb_GN;
    } else {
        int t = tok - sub_len - 1;
        int row = t / 13, col = t % 13;
        if (col == 12) gn = sub_GN;
        else { crop = 0; r = row; c = col; }
    }
    for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
        if (gn) { o[i] = gn[i]; continue; }
        int q = i / C, ch = i % C;
        int fr = 2 * r + (q >> 1), fc = 2 * c + (q & 1);
        o[i] = __float2bfloat16_rn(feats[((size_t)crop * 577 + 1 + fr * 24 + fc) * C + ch]);
    }
}

extern "C" int p3_gn_assemble(const float* feats, const void* sub_GN, const void* glb_GN, void* out, int hc, int wc,
                              int C, cudaStream_t st) {
    P3_CHECK_ARG(hc >= 1 && wc >= 1, "gn_assemble: bad crop grid");
    int total_tok = hc * wc * 144 + 12 * hc + 1 + 156;
    gn_assemble_kernel<<<total_tok, 256, 0, st>>>(feats, (const bf16*)sub_GN, (const bf16*)glb_GN, (bf16*)out, hc, wc, C);
    P3_CHECK_LAUNCH("gn_assemble");
    return 0;
}
