"""4-bit weight quantisation of `load(quantize_model=True)` (reference: `nn.quantize(model, 64, 4)`,
phi_3_vision_mlx.py:264,291-305 — every nn.Linear / nn.Embedding weight, group 64 along the input
dimension, affine codes). Product-side torch code (runs on the GPU at load time); the CPU oracle
keeps its own restatement (oracle/phi3_oracle.py::quantize_model_weights) and the tests compare the two.

Rule [mlx-inferred, same as the KV quantiser in csrc/kvquant.cu]: per group, scale = (max-min)/15 with the
sign chosen so that the larger-magnitude edge is exactly representable and 0 stays representable; scale and
bias are stored as bf16 (the checkpoint dtype) and the codes are computed against the stored values:
q = clip(round((w - bias) / scale), 0, 15);  dequantised weight = q * scale + bias.

Two images of a quantised matrix are kept on the device:
  * `deq`   bf16 [N, K] = bf16(q*scale + bias): operand of the tcgen05 GEMMs (prefill, ViT, projector);
  * `codes` uint8 [N, K/2] + `meta` bf16 [N, K/64, 2]: the 4-bit stream of the decode-time skinny GEMM
    (csrc/gemm_skinny.cu, W4 variant), packed in the lane order that kernel reads (see pack_w4g64).
"""
import torch

GROUP = 64
_KOFF = (0, 2, 4, 6, 1, 3, 5, 7)      # nibble p of a 32-bit word holds k-offset _KOFF[p] of its 8 weights


def quantize_w4g64(w):
    """w [N, K] (any float dtype) -> codes uint8 [N, K] in 0..15, scale bf16 [N, K/64], bias bf16 [N, K/64]."""
    N, K = w.shape
    assert K % GROUP == 0, f'quantize_model: input dimension {K} is not a multiple of {GROUP}'
    # float64 throughout: IEEE division and rounding give the same codes on every device (in fp32 a quotient that lands
    # within an ulp of k + 0.5 rounds differently on CPU and GPU, which moves that weight by a whole quantisation step)
    g = w.reshape(N, K // GROUP, GROUP).to(torch.float64)
    w_max, w_min = g.amax(-1, keepdim=True), g.amin(-1, keepdim=True)
    neg = w_min.abs() > w_max.abs()
    # tensor divisor on purpose: torch's CUDA kernel turns `x / python_scalar` into `x * (1 / scalar)`, which is not the
    # IEEE quotient the CPU computes and flips ties in the edge rule below
    scale = torch.clamp((w_max - w_min) / torch.full_like(w_max, 15.0), min=1e-7)
    scale = torch.where(neg, scale, -scale)
    edge = torch.where(neg, w_min, w_max)
    q0 = torch.round(edge / scale)
    scale = torch.where(q0 != 0, edge / q0, scale)
    bias = torch.where(q0 == 0, torch.zeros_like(edge), edge)
    scale = scale.to(torch.float32).to(torch.bfloat16)          # explicit two-step rounding (what a device does in one
    bias = bias.to(torch.float32).to(torch.bfloat16)            # fp64 -> bf16 cast is not the same everywhere)
    sf = scale.to(torch.float64)
    sf = torch.where(sf == 0, torch.full_like(sf, 1e-7), sf)
    codes = torch.clamp(torch.round((g - bias.to(torch.float64)) / sf), 0, 15).to(torch.uint8)
    return codes.reshape(N, K), scale.reshape(N, K // GROUP), bias.reshape(N, K // GROUP)


def dequantize_w4g64(codes, scale, bias):
    """-> bf16 [N, K] = bf16(q * scale + bias), computed in fp32."""
    N, K = codes.shape
    q = codes.reshape(N, K // GROUP, GROUP).to(torch.float32)
    w = q * scale.to(torch.float32)[..., None] + bias.to(torch.float32)[..., None]
    return w.reshape(N, K).to(torch.bfloat16)


def pack_w4g64(codes, scale, bias):
    """Kernel layout. Per row and per 128-wide super-chunk c: 64 bytes = 4 lanes x 16 bytes; lane t's 16 bytes are
    [group 2c: word 0, word 1][group 2c+1: word 0, word 1]; word i of lane t covers k = 64*grp + 16*t + 8*i + (0..7)
    with nibble p (bits 4p..4p+3) holding k-offset (0,2,4,6,1,3,5,7)[p], so that `(word >> 4j) & 0x000F000F` yields the
    adjacent pair (k=2j, k=2j+1) as two 16-bit lanes. meta[n, grp] = (scale, bias) bf16."""
    N, K = codes.shape
    assert K % 128 == 0, f'W4 skinny GEMM needs K % 128 == 0 (got {K})'
    c = codes.reshape(N, K // 128, 2, 4, 2, 8)                     # [n, chunk, grp, lane, word, k-offset]
    c = c[..., list(_KOFF)]                                        # nibble order
    c = c.permute(0, 1, 3, 2, 4, 5).contiguous().to(torch.int64)   # [n, chunk, lane, grp, word, nibble]
    shifts = (4 * torch.arange(8, device=codes.device, dtype=torch.int64))
    words = (c << shifts).sum(-1)                                  # [n, chunk, lane, grp, word] uint32 values in int64
    b = torch.stack([(words >> (8 * i)) & 0xFF for i in range(4)], -1).to(torch.uint8)   # little-endian bytes
    packed = b.reshape(N, K // 2).contiguous()
    meta = torch.stack([scale, bias], -1).to(torch.bfloat16).contiguous()                # [N, K/64, 2]
    return packed, meta


class W4:
    """A quantised [N, K] matrix on the device (see module docstring)."""

    def __init__(self, w, pack=True, row_perm=None):
        codes, scale, bias = quantize_w4g64(w)
        if row_perm is not None:                                    # e.g. the gate/up interleave of the SwiGLU epilogue
            codes, scale, bias = row_perm(codes), row_perm(scale), row_perm(bias)
        self.deq = dequantize_w4g64(codes, scale, bias)
        self.codes, self.meta = pack_w4g64(codes, scale, bias) if pack else (None, None)
        self.shape = tuple(self.deq.shape)


def fake_quant(w):
    """bf16(dequantise(quantise(w))) for weights that only ever feed the tensor-core GEMMs."""
    shp = w.shape
    w2 = w.reshape(-1, shp[-1])
    return dequantize_w4g64(*quantize_w4g64(w2)).reshape(shp)
