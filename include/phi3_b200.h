/* phi3_b200.h — C ABI of libphi3b200.so (sm_100a only).
 *
 * The reference (JosefAlbers/Phi-3-Vision-MLX) has no FFI/plugin layer: its hot path calls
 * mlx==0.15 ops from Python. Each entry below replaces one of those MLX call
 * sites; the cited lines are in /root/reference/phi.py (phi:) and
 * /root/reference/phi_3_vision_mlx.py (pv:). INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions: plain device pointers + sizes, caller-owned memory and workspaces, no
 * allocation and no host synchronisation inside, stream-ordered and CUDA-graph capturable.
 * Every function returns 0 on success, <0 on error; p3_last_error() gives the message
 * (thread-local). bf16 = __nv_bfloat16 storage. "Rows" are tokens (B*L flattened).
 */
#ifndef PHI3_B200_H
#define PHI3_B200_H
#include <stdint.h>
#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

/* GEMM epilogues (p3_gemm, p3_gemm_skinny) */
#define P3_EPI_NONE 0       /* out bf16 = acc (+bias) */
#define P3_EPI_QGELU 1      /* bias + x*sigmoid(1.702x)         nn.gelu_fast_approx, phi:154 */
#define P3_EPI_GELU 2       /* bias + exact erf GELU            nn.GELU(), phi:391 */
#define P3_EPI_RESIDUAL 3   /* out = resid + bf16(acc+bias)     phi:170-171, 483, 485 */
#define P3_EPI_SWIGLU 4     /* out[:, N/2] = silu(gate)*up      phi:470-471, interleaved W (see p3_gemm) */
#define P3_EPI_F32 5        /* out fp32 = acc (+bias)           lm_head logits, phi:608 */
#define P3_EPI_RESIDUAL_F32 6 /* out fp32 = resid fp32 + acc+bias  CLIP residual stream (fp32 in the reference) */

#define P3_PAGE_TOKENS 64   /* tokens per KV page */

const char* p3_last_error(void);
int p3_version(void);
/* Diagnostics (tools/chain_trace.py), not part of the reference interface: with a device buffer of n_slots x 1024 x 8 uint64 set,
 * every decode-path launch (skinny GEMMs, decode attention) takes the next slot and its CTAs store globaltimer stamps
 * {start, dependency resolved, first operands ready, main loop done, exit, smid, kind, grid}. p3_trace_set(NULL, 0) turns it off. */
int p3_trace_set(void* buf, int n_slots);
int p3_trace_count(void);

/* nn.Embedding phi:568,577 — ids<0 (image placeholders, phi:270) read row 0. ss_out (fp32 [T], may be
 * NULL) receives each row's sum of squares for the RMSNorm fused into the next p3_gemm_skinny. */
int p3_embed_gather(const void* table, const int32_t* ids, void* out, int64_t T, int H, int vocab, float* ss_out,
                    cudaStream_t st);

/* nn.RMSNorm / mx.fast.rms_norm phi:478-479,571 */
int p3_rmsnorm(const void* x, const void* w, void* y, int64_t T, int H, float eps, cudaStream_t st);

/* nn.LayerNorm / mx.fast.layer_norm phi:165,167,212. x fp32 (CLIP residual stream);
 * y bf16 (out_f32 = 0, feeds a GEMM) or fp32 (out_f32 = 1, pre_layrnorm phi:218). */
int p3_layernorm(const float* x, const void* w, const void* b, void* y, int64_t T, int H, float eps, int out_f32,
                 cudaStream_t st);

/* _rotate_half + SuRoPE table + KVCache slice-assign: phi:418-423, 487-507, 542-548.
 * qkv [B*L, (n_heads+2*n_kv)*hd] bf16 roped in place; cos/sin fp32 [Bt, L_all, hd/2];
 * tab_bstride = 0 for a shared table, else elements between table rows; row b uses
 * table/cache row b/row_div (row_div = n_beam, phi:447-450). pool = this layer's page pool
 * [page][2][n_kv][64][hd]; block_table int32 [n_seq, bt_stride]. past_dev (device int32, may be
 * NULL) overrides `past` at run time so a captured CUDA graph can be replayed every step. */
int p3_rope_kvwrite(void* qkv, const float* cosT, const float* sinT, int64_t tab_bstride, int B, int L, int n_heads,
                    int n_kv, int hd, int past, int row_div, void* pool, const int32_t* block_table, int bt_stride,
                    int write_cache, const int32_t* past_dev, cudaStream_t st);

/* mx.argmax pv:386,392,506,559; nn.log_softmax + gathers pv:476,541-547,571-573;
 * mx.argpartition top-n pv:507 — one pass per logits row. Any output pointer may be NULL. */
int p3_row_stats(const float* logits, int64_t R, int64_t ld, int V, int32_t* argmax_out, float* max_out,
                 float* lse_out, int n_top, int32_t* topk_ids, float* topk_lp, int n_gather,
                 const int32_t* gather_ids, float* gather_lp, cudaStream_t st);

/* Nucleus (top-p) sampling — extension: the reference decodes greedily only (pv:386,392).
 * tau = largest probability v with sum_{p_i>=v} p_i >= top_p (exact bisection, no sort); token drawn by
 * inverse CDF in index order over {p_i >= tau} with the caller's uniform u[R] in [0,1). tau_out may be NULL.
 * step_dev (device int32, may be NULL): u is a table and row *step_dev * u_stride is used (graph replay). */
int p3_top_p_sample(const float* logits, int64_t R, int64_t ld, int V, float top_p, float temperature, const float* u,
                    int32_t* out, float* tau_out, const int32_t* step_dev, int64_t u_stride, cudaStream_t st);

/* Greedy-loop bookkeeping on the device (pv:390-398 without the two host syncs per token):
 * history[b][*step] = tok[b]; eos_seen[b] |= (tok[b]==32007); ++*step; ++*past (past may be NULL). */
int p3_decode_advance(const int32_t* tok, int32_t* history, int64_t ld, int B, int32_t* step, int32_t* past,
                      int32_t* eos_seen, cudaStream_t st);

/* nn.Linear at decode (M<=16 tokens): phi:437-438,465-466,604. Optional fused RMSNorm prologue
 * (norm_w != NULL: X is the raw hidden state). epi in {NONE, RESIDUAL, SWIGLU, F32}.
 * ss_in (fp32 [n_ss_in][16], may be NULL): per-token partial sums of squares of X written by the
 * kernel that produced X (p3_embed_gather / a RESIDUAL p3_gemm_skinny via ss_out, which writes
 * [ceil(N/16)][16]); without it the RMSNorm statistic is recomputed from X.
 * l2_prefetch (may be NULL): the NEXT kernel's weights; every CTA pulls a slice into L2 before it waits on its
 * predecessor, so HBM keeps streaming across the kernel boundary (own weights are loaded L2::evict_first). */
int p3_gemm_skinny(const void* X, int64_t ldx, const void* norm_w, float eps, const void* W, void* out, int64_t ldo,
                   const void* resid, int M, int N, int K, int epi, const float* ss_in, int n_ss_in, float* ss_out,
                   const void* l2_prefetch, int64_t l2_prefetch_bytes, cudaStream_t st);

/* Decode-time qkv_proj + _rotate_half/SuRoPE + KVCache write in ONE launch (phi:442-453): each CTA owns
 * 16 rotary pairs (column c and c + hd/2 of one head) so the rotation fuses into the epilogue. */
int p3_gemm_skinny_qkv_rope(const void* X, int64_t ldx, const void* norm_w, float eps, const void* Wqkv, void* qkv,
                            const float* ss_in, int n_ss_in, const float* cosT, const float* sinT, int64_t tab_bstride,
                            int B, int L, int n_heads, int n_kv, int hd, int K, int past, const int32_t* past_dev,
                            int row_div, void* pool, const int32_t* block_table, int bt_stride, int write_cache,
                            const void* l2_prefetch, int64_t l2_prefetch_bytes, cudaStream_t st);

/* quantize_model=True (pv:264,291-305: nn.quantize(model, group_size=64, bits=4) -> QuantizedLinear): the two decode-time
 * entries above over the 4-bit image of W. Wq uint8 [N][K/2] and Wmeta bf16 [N][K/64][2] = (scale, bias) in the lane order
 * written by phi3_b200/quant.py::pack_w4g64; y = sum_k x_k (scale_g q_k + bias_g) accumulated in fp32 (codes and a
 * matrix of ones go through the tensor cores per 64-wide group, the affine map is applied to the group sums). K % 128 == 0.
 * Prefill / ViT GEMMs of a quantised model run p3_gemm on the bf16 dequantised image. */
int p3_gemm_skinny_w4(const void* X, int64_t ldx, const void* norm_w, float eps, const void* Wq, const void* Wmeta, void* out,
                      int64_t ldo, const void* resid, int M, int N, int K, int epi, const float* ss_in, int n_ss_in,
                      float* ss_out, const void* l2_prefetch, int64_t l2_prefetch_bytes, cudaStream_t st);
int p3_gemm_skinny_qkv_rope_w4(const void* X, int64_t ldx, const void* norm_w, float eps, const void* Wq, const void* Wmeta,
                               void* qkv, const float* ss_in, int n_ss_in, const float* cosT, const float* sinT,
                               int64_t tab_bstride, int B, int L, int n_heads, int n_kv, int hd, int K, int past,
                               const int32_t* past_dev, int row_div, void* pool, const int32_t* block_table, int bt_stride,
                               int write_cache, const void* l2_prefetch, int64_t l2_prefetch_bytes, cudaStream_t st);

/* nn.Linear for prefill / ViT / projector (phi:140-143,155-156,391,437-438,465-466,604) and the
 * patch-embed conv as GEMM (phi:186-192): out[M,N] = X[M,K] . W[N,K]^T on tcgen05 tensor cores
 * (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue). bias bf16 [N] or NULL. row_map int32 [M] or
 * NULL scatters output rows (image-feature splice, phi:412-415). For P3_EPI_SWIGLU W must be in
 * the interleaved layout (per 256 rows: 128 gate rows then the matching 128 up rows) and out is
 * [M, N/2]. impl: 0 = tcgen05 (product path), 1 = mma.sync cross-check kernel (tests only). */
int p3_gemm(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* out, int64_t ldo,
            const void* resid, const int32_t* row_map, int64_t M, int N, int K, int epi, int impl, cudaStream_t st);

/* Prefill / ViT flash attention: phi:454-457 (causal + left-pad predicate instead of Mask4D
 * phi:550-563) and mx.fast.scaled_dot_product_attention phi:148 (causal=0).
 * q/k/v point into the (roped) qkv buffer; row strides in elements; head h at +h*hd.
 * Keys [0,past) come from the paged pool (cache row b/row_div), keys [past,past+L) from k/v.
 * kv_start int32 [B] (left-pad length; NULL = 0). Pad queries produce zeros.
 * Runs the tcgen05 flash-attention kernel (S and P in TMEM, TMA operands) when past % 64 == 0, L >= 64 and
 * the pointers / strides are 16-byte aligned; otherwise (and with P3_ATTN_TC=0 in the environment, used by the
 * tests as a cross-check) the mma.sync kernel with the same contract. */
int p3_attention_prefill(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* out,
                         int64_t ldo, int B, int L, int n_heads, int n_kv, int hd, float scale, int causal, int past,
                         const int32_t* kv_start, const void* pool, const int32_t* block_table, int bt_stride,
                         int row_div, cudaStream_t st);

/* Decode-time attention for L<=16 new tokens per row over a paged KV cache (split-KV):
 * phi:454-457 with KVCache reads phi:523-527 (n_beam shared prefix: row_div) / phi:548.
 * workspace: p3_attention_decode_workspace() bytes, MUST be zero-filled once before first use (it holds the
 * split-arrival counters, which the kernel resets itself); needed when n_splits > 1.
 * l2_prefetch (may be NULL): l2_prefetch_bytes of the NEXT kernel's weights (o_proj) that the CTAs pull into
 * L2 (prefetch.global.L2::evict_last) while the KV pages stream. */
/* Struct-argument form of p3_gemm_skinny / p3_gemm_skinny_w4 / p3_gemm_skinny_qkv_rope(_w4) with the RMSNorm split between
 * producer and consumer (decode over 4-bit weights, phi:478-485 with QuantizedLinear pv:264): a RESIDUAL launch with xg_out also
 * writes xg_out[m][n] = bf16(h[m][n] * xg_gain[n]) (xg_gain = weight of the NEXT RMSNorm); the consumer takes that as X with
 * norm_w = NULL and rs_epi = 1 and multiplies its reduced accumulators by rsqrt(sum(ss_in[.][m]) / K + eps). */
typedef struct {
    int32_t op;                          /* 0: out = epi(X W^T);  1: qkv projection + SuRoPE + paged KV write */
    const void* X; int64_t ldx;
    const void* norm_w; float eps;
    const void* W;                       /* bf16 [N][K], or NULL when Wq / Wmeta (4-bit g64 image) are given */
    const void* Wq; const void* Wmeta;
    void* out; int64_t ldo; const void* resid;
    int32_t M, N, K, epi;                /* op 1: N, M, ldo, epi are derived (M = B * L) */
    const float* ss_in; int32_t n_ss_in; float* ss_out;
    const void* l2_prefetch; int64_t l2_prefetch_bytes;
    const void* xg_gain; void* xg_out; int64_t ldxg; int32_t rs_epi;
    const float* cosT; const float* sinT; int64_t tab_bstride;          /* op 1, as p3_gemm_skinny_qkv_rope */
    int32_t B, L, n_heads, n_kv, hd, past, row_div, bt_stride, write_cache;
    const int32_t* past_dev; void* pool; const int32_t* block_table;
    int32_t packed;                      /* op 0, bf16 W, N % 16 == 0, N < 9472, not SwiGLU: W is given in stream order
                                          * [N/16 tiles][K/64 chunks][row half][k half][lane = 4 * (row % 8) + 16-byte quad][8 bf16]
                                          * (one contiguous 16 x K block per tile; phi3_b200.model.pack_rows16) */
} p3_skinny_args;
int p3_gemm_skinny_x(const p3_skinny_args* args, cudaStream_t st);
/* p3_embed_gather that also writes xg_out[t] = bf16(row * xg_gain) (input of the first rs_epi consumer) */
int p3_embed_gather_xg(const void* table, const int32_t* ids, void* out, int64_t T, int H, int vocab, float* ss_out,
                       const void* xg_gain, void* xg_out, cudaStream_t st);

int64_t p3_attention_decode_workspace(int B, int L, int n_heads, int hd, int n_splits);
int p3_attention_decode(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* out,
                        int64_t ldo, int B, int L, int n_heads, int n_kv, int hd, float scale, int past,
                        const int32_t* kv_start, const void* pool, const int32_t* block_table, int bt_stride,
                        int row_div, int n_splits, void* workspace, const int32_t* past_dev, const void* l2_prefetch,
                        int64_t l2_prefetch_bytes, cudaStream_t st);

/* 4-bit g32 KV-cache quantisation (mx.quantize/dequantize, phi:532,536-537). Quantises the
 * first n_tokens positions of the bf16 pool pages of each cache row in place into a q4 pool
 * ([page][2][n_kv][64][hd/2 bytes] codes + [page][2][n_kv][64][hd/32][2] bf16 scale,bias). */
int p3_kv_quantize_q4g32(const void* pool, void* qcodes, void* qmeta, const int32_t* block_table, int bt_stride,
                         int n_seq, int n_tokens, int n_kv, int hd, cudaStream_t st);
int p3_attention_decode_q4(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                           void* out, int64_t ldo, int B, int L, int n_heads, int n_kv, int hd, float scale, int past,
                           int n_quant, const int32_t* kv_start, const void* pool, const void* qcodes,
                           const void* qmeta, const int32_t* block_table, int bt_stride, int row_div, int n_splits,
                           void* workspace, const int32_t* past_dev, const void* l2_prefetch, int64_t l2_prefetch_bytes,
                           cudaStream_t st);

/* Vision front end. */
/* PIL BILINEAR resize (phi:299), white pad to x336 (phi:300-306), (x/255-mean)/std (phi:309):
 * src uint8 HWC with element strides (sx, sy) so a transposed view needs no copy (phi:295,308).
 * coefficient tables are built on the host exactly as Pillow's precompute_coeffs does. */
int p3_hd_resize_h(const uint8_t* src, int64_t sy, int64_t sx, int in_w, int in_h, uint8_t* tmp, int out_w,
                   const int32_t* bounds, const int32_t* kk, int ksize, cudaStream_t st);
/* vertical pass + white pad (pad_top rows above, to padded_h) + optional un-transpose; writes the
 * padded uint8 HWC image ([padded_h, tmp_w, 3], or [tmp_w, padded_h, 3] when transposed). */
int p3_hd_resize_v_pad(const uint8_t* tmp, int tmp_w, int tmp_h, int out_h, const int32_t* bounds, const int32_t* kk,
                       int ksize, int pad_top, int padded_h, int transposed, uint8_t* out_hwc, cudaStream_t st);
/* normalise through a 256x3 float64 LUT ((v/255-mean)/std, phi:309), sub-crop tiling (phi:322-326)
 * and the reference's 2-tap interpolate_336 global crop (phi:331-372, float64 accumulate) into
 * pixel_values [n_crops+1, 3, 336, 336] fp32 (crop 0 = global). idx/wgt tables are [336][2]. */
int p3_hd_tile_crops(const uint8_t* img_hwc, int H, int W, const double* lut, float* pixel_values,
                     const int32_t* h_idx, const float* h_wgt, const int32_t* w_idx, const float* w_wgt,
                     cudaStream_t st);
/* im2col for the 14x14/14 patch conv (phi:186-192,199): pixel_values [N,3,336,336] fp32 ->
 * A [N*576, Kpad] bf16, k = (ky*14+kx)*3 + c (NHWC weight order, pv:374), zero padded. */
int p3_patch_im2col(const float* pixel_values, void* A, int N, int Kpad, cudaStream_t st);
/* cat[class_emb, patches] + position_embedding (phi:202-205): patches fp32 [N*576, D] -> fp32 [N,577,D] */
int p3_clip_embed(const float* patches, const void* cls, const void* pos, float* out, int N, int D, cudaStream_t st);
/* drop CLS (phi:221) + 2x2 token merge + sub_GN/glb_GN separators (phi:403-407):
 * feats fp32 [n_crops+1, 577, C] (crop 0 = global) -> bf16 [(hc*wc+1)*144 + 1 + (hc+1)*12, 4*C] */
int p3_gn_assemble(const float* feats, const void* sub_GN, const void* glb_GN, void* out, int hc, int wc, int C,
                   cudaStream_t st);

/* ------------------------------------------------------------------------------------------------
 * Struct-argument GEMM entry with the fused prefill epilogues (north_star: "SuRoPE fused into the QKV epilogue",
 * "RMSNorm and residual fused with adjacent ops"); replaces nn.RMSNorm + nn.Linear + _rotate_half + KVCache slice-assign at
 * prefill: phi:442-453, 478-485.
 *   - ss_in / n_ss_in / eps: RMSNorm of the INPUT rows is applied as a per-row scale rsqrt(sum_c ss_in[row][c] / K + eps) on the
 *     accumulators; the gain must already be folded into W (W'[n][k] = W[n][k] * g[k]). ss_in rows hold partial sums of
 *     squares of X's rows (p3_row_sumsq, or the ss_out of the GEMM that produced X).
 *   - ss_out (P3_EPI_RESIDUAL): fp32 [M][N/32], sum of squares of every 32-column chunk written.
 *   - P3_EPI_ROPE_KV: out = qkv [M, N] bf16 with q,k rotated, K/V also written to the page pool (as p3_rope_kvwrite). The q
 *     and k rows of W must be permuted per head to [16j..16j+15 | half+16j..half+16j+15], j = 0..hd/32-1 (so a rotary pair
 *     meets in one 32-column chunk of the accumulator); v rows keep their order.
 *   - splitk_ws: see the field comment (deterministic split-K: partials are summed in slice order by the last CTA of a tile).
 *   - w_plan: NULL, or a P3_GEMM_WPLAN_BYTES blob filled once by p3_gemm_plan_weights for this W (skips re-encoding the
 *     weight tensor map on every call). */
#define P3_EPI_ROPE_KV 8
#define P3_GEMM_WPLAN_BYTES 320
typedef struct {
    const void* X; int64_t ldx;
    const void* W; int64_t ldw;
    const void* bias; void* out; int64_t ldo;
    const void* resid; const int32_t* row_map;
    int64_t M; int32_t N, K, epi, impl;
    const float* ss_in; int32_t n_ss_in; float eps;
    float* ss_out;
    const void* w_plan;
    /* P3_EPI_ROPE_KV (M = B*L rows, token i of row b at position past+i; table / cache row b/row_div) */
    const float* cosT; const float* sinT; int64_t tab_bstride;
    int32_t L, n_heads, n_kv, hd, past, row_div, write_cache, bt_stride;
    const int32_t* past_dev;
    void* pool; const int32_t* block_table;
    /* split-K for shapes with few output tiles (M <= 128 ...): NULL, or a caller-owned device workspace whose first 4 KB are
     * zero before the first use (per-tile arrival counters, left at zero by every launch) followed by room for the fp32
     * partial tiles (32 MB covers Phi-3.5 shapes). Without it such shapes run on as many SMs as they have tiles. */
    void* splitk_ws; int64_t splitk_ws_bytes;
} p3_gemm_args;
int p3_gemm_plan_weights(const void* W, int64_t ldw, int N, int K, void* plan);
int p3_gemm_fused(const p3_gemm_args* args, cudaStream_t st);
/* SuRoPE cos/sin table phi:487-507 on the device: cosT/sinT fp32 [Bt, L_all, half]; positions 0..L_all-1 (pids NULL) or per row
 * cat[pids[b, 0..Lp), pids[b, Lp-1] + 1 + arange] (phi:493-497); inv_freq fp32 [half] = 1 / (factor * theta^(2i/dim)). */
int p3_rope_table(const int32_t* pids, int64_t pid_stride, int Lp, const float* inv_freq, float* cosT, float* sinT, int Bt,
                  int L_all, int half, float scale, cudaStream_t st);
/* fp32 out[row] = sum of squares of bf16 x[row, 0..H) (feeds ss_in with n_ss_in = 1) */
int p3_row_sumsq(const void* x, int64_t ldx, float* out, int64_t T, int H, cudaStream_t st);

/* ------------------------------------------------------------------------------------------------
 * Persistent decode-layer kernel (decode_mega.cu): for M <= 8 decode rows, ONE launch of n_ctas (<= #SM,
 * co-resident) persistent CTAs runs up to 4 chained weight-stream phases with grid barriers in between,
 * e.g. o_proj(+residual) -> RMSNorm+gate_up(+SwiGLU) -> down_proj(+residual) -> RMSNorm+qkv_proj(+SuRoPE
 * +paged KV write) of the next layer, or ... -> RMSNorm+lm_head. Replaces nn.Linear / nn.RMSNorm /
 * _rotate_half / KVCache slice-assign at decode: phi:437-438, 442-453, 460, 465-471, 478-485, 604-608.
 * Arguments travel in a plain-old-data struct (device pointers + sizes), as SURVEY.md par. 8b asks. */
#define P3_MEGA_RESID 0     /* out[M,N] bf16 = bf16(out + bf16(x W^T)) in place; ss_out[tile][16] = sum of squares of the 16 new columns */
#define P3_MEGA_SWIGLU 1    /* W = [gate rows | up rows] (phi:470); out[M,N/2] bf16 = silu(gate) * up */
#define P3_MEGA_QKV_ROPE 2  /* out[M,N] bf16 = qkv with q,k rotated (phi:418-423) + K,V written to the page pool at `past` */
#define P3_MEGA_F32 3       /* out[M,N] fp32 logits */
#define P3_MEGA_MAX_PHASES 4

typedef struct {
    const void* wp;             /* weights in stream order (p3_mega_pack), N*K bf16 */
    int32_t kind, N, K;         /* N = rows of W (output features, SWIGLU: gate+up), K = input features */
    int32_t max_tiles_per_cta;  /* max over CTAs of the schedule below (checked against the smem partial buffer when K > 4096) */
    const void* x;              /* input activations bf16 [M, ldx] */
    int64_t ldx;
    const void* norm_w;         /* RMSNorm gain applied to x on load (bf16 [K]) or NULL */
    const float* ss_in;         /* per-row sum-of-squares partials of x: fp32 [n_ss_in][16]; NULL = recompute from x */
    int32_t n_ss_in, _pad;
    float* ss_out;              /* RESID: partials [N/16][16] of the rows written, or NULL */
    void* out;
    int64_t ldo;
    const int32_t* cta_off;     /* schedule: CTA c owns tile_ids[cta_off[c] .. cta_off[c+1]) of this phase (int32 [n_ctas+1]) */
    const int32_t* tile_ids;    /* tile = 16 (RESID) or 2x16 (other kinds) rows of W */
} p3_mega_phase;

typedef struct {
    p3_mega_phase ph[P3_MEGA_MAX_PHASES];
    int32_t n_phases, M;
    float eps;                  /* RMSNorm eps */
    int32_t n_ctas;             /* grid size the schedule was built for (<= #SM) */
    uint32_t* sync;             /* 2 words of device memory, zero before the first launch (grid barrier state) */
    /* P3_MEGA_QKV_ROPE: one new token per row at position *past_dev (or `past` when NULL) */
    const float* cosT;
    const float* sinT;          /* fp32 [Bt, L_all, hd/2]; row b uses table row b (tab_bstride elements apart, 0 = shared) */
    int64_t tab_bstride;
    int32_t n_heads, n_kv, hd, past;
    const int32_t* past_dev;
    void* pool;                 /* this layer's page pool [page][2][n_kv][64][hd] bf16 */
    const int32_t* block_table; /* int32 [M, bt_stride] */
    int32_t bt_stride, _pad2;
    long long* dbg;             /* NULL, or int64 [n_ctas][4][8]: per-CTA / per-phase SM-clock stamps (tools/mega_trace.py) */
} p3_mega_args;

/* W [N,K] bf16 (nn.Linear layout; SWIGLU: the checkpoint's [gate | up] row order) -> stream order for `kind` */
int p3_mega_pack(const void* W, void* out, int kind, int N, int K, int n_heads, int n_kv, int hd, cudaStream_t st);
/* number of SMs of the current device = the largest valid n_ctas */
int p3_decode_mega_ctas(void);
int p3_decode_mega(const p3_mega_args* args, cudaStream_t st);

#ifdef __cplusplus
}
#endif
#endif
