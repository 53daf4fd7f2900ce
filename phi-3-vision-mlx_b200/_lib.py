"""ctypes binding of libphi3b200.so (the C ABI declared in include/phi3_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, this
module raises. Build with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C phi-3-vision-mlx_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('P3_LIB') or os.path.join(_HERE, 'libphi3b200.so')     # P3_LIB: tuning builds (tools/), same ABI

EPI_NONE, EPI_QGELU, EPI_GELU, EPI_RESIDUAL, EPI_SWIGLU, EPI_F32, EPI_RESIDUAL_F32 = range(7)
EPI_ROPE_KV = 8
GEMM_WPLAN_BYTES = 320
PAGE = 64

_p, _i, _l, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_SIGS = {
    'p3_embed_gather': [_p, _p, _p, _l, _i, _i, _p, _p],
    'p3_rmsnorm': [_p, _p, _p, _l, _i, _f, _p],
    'p3_layernorm': [_p, _p, _p, _p, _l, _i, _f, _i, _p],
    'p3_rope_kvwrite': [_p, _p, _p, _l, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p],
    'p3_row_stats': [_p, _l, _l, _i, _p, _p, _p, _i, _p, _p, _i, _p, _p, _p],
    'p3_gemm_skinny': [_p, _l, _p, _f, _p, _p, _l, _p, _i, _i, _i, _i, _p, _i, _p, _p, _l, _p],
    'p3_gemm_skinny_qkv_rope': [_p, _l, _p, _f, _p, _p, _p, _i, _p, _p, _l, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _i, _p, _l, _p],
    'p3_gemm_skinny_w4': [_p, _l, _p, _f, _p, _p, _p, _l, _p, _i, _i, _i, _i, _p, _i, _p, _p, _l, _p],
    'p3_gemm_skinny_qkv_rope_w4': [_p, _l, _p, _f, _p, _p, _p, _p, _i, _p, _p, _l, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _i, _p, _l, _p],
    'p3_gemm': [_p, _l, _p, _l, _p, _p, _l, _p, _p, _l, _i, _i, _i, _i, _p],
    'p3_attention_prefill': [_p, _p, _p, _l, _l, _l, _p, _l, _i, _i, _i, _i, _i, _f, _i, _i, _p, _p, _p, _i, _i, _p],
    'p3_attention_decode': [_p, _p, _p, _l, _l, _l, _p, _l, _i, _i, _i, _i, _i, _f, _i, _p, _p, _p, _i, _i, _i, _p, _p, _p, _l, _p],
    'p3_attention_decode_q4': [_p, _p, _p, _l, _l, _l, _p, _l, _i, _i, _i, _i, _i, _f, _i, _i, _p, _p, _p, _p, _p, _i,
                               _i, _i, _p, _p, _p, _l, _p],
    'p3_kv_quantize_q4g32': [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    'p3_top_p_sample': [_p, _l, _l, _i, _f, _f, _p, _p, _p, _p, _l, _p],
    'p3_decode_advance': [_p, _p, _l, _i, _p, _p, _p, _p],
    'p3_hd_resize_h': [_p, _l, _l, _i, _i, _p, _i, _p, _p, _i, _p],
    'p3_hd_resize_v_pad': [_p, _i, _i, _i, _p, _p, _i, _i, _i, _i, _p, _p],
    'p3_hd_tile_crops': [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    'p3_patch_im2col': [_p, _p, _i, _i, _p],
    'p3_clip_embed': [_p, _p, _p, _p, _i, _i, _p],
    'p3_gn_assemble': [_p, _p, _p, _p, _i, _i, _i, _p],
    'p3_mega_pack': [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    'p3_row_sumsq': [_p, _l, _p, _l, _i, _p],
    'p3_rope_table': [_p, _l, _i, _p, _p, _p, _i, _i, _i, _f, _p],
    'p3_embed_gather_xg': [_p, _p, _p, _l, _i, _i, _p, _p, _p, _p],
}

_lib = None
launches = 0          # number of p3_* kernel-launching calls issued (bench.py reports it)


def exported_symbols():
    return sorted(list(_SIGS) + ['p3_last_error', 'p3_version', 'p3_attention_decode_workspace', 'p3_decode_mega',
                                 'p3_decode_mega_ctas', 'p3_gemm_fused', 'p3_gemm_plan_weights', 'p3_gemm_skinny_x', 'p3_trace_set', 'p3_trace_count'])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: the CUDA extension must be built (no CPU fallback exists). '
                               'Run __graft_entry__.build().')
        L = C.CDLL(LIB_PATH)
        for name, sig in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = sig, C.c_int
        L.p3_last_error.restype = C.c_char_p
        L.p3_version.restype = C.c_int
        L.p3_attention_decode_workspace.argtypes = [_i, _i, _i, _i, _i]
        L.p3_attention_decode_workspace.restype = C.c_int64
        L.p3_decode_mega.argtypes, L.p3_decode_mega.restype = [_p, _p], C.c_int      # (const p3_mega_args*, stream)
        L.p3_decode_mega_ctas.argtypes, L.p3_decode_mega_ctas.restype = [], C.c_int
        L.p3_gemm_fused.argtypes, L.p3_gemm_fused.restype = [_p, _p], C.c_int        # (const p3_gemm_args*, stream)
        L.p3_gemm_plan_weights.argtypes, L.p3_gemm_plan_weights.restype = [_p, _l, _i, _i, _p], C.c_int
        L.p3_gemm_skinny_x.argtypes, L.p3_gemm_skinny_x.restype = [_p, _p], C.c_int        # (const p3_skinny_args*, stream)
        L.p3_trace_set.argtypes, L.p3_trace_set.restype = [_p, C.c_int], C.c_int           # diagnostics (tools/chain_trace.py)
        L.p3_trace_count.argtypes, L.p3_trace_count.restype = [], C.c_int
        _lib = L
    return _lib


def ptr(t):
    """Device (or host) address of a torch tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


def call(name, *args):
    global launches
    L = lib()
    if len(args) != len(_SIGS[name]):
        raise TypeError(f'{name}: expected {len(_SIGS[name])} arguments, got {len(args)}')
    rc = getattr(L, name)(*args)
    launches += 1
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {L.p3_last_error().decode()}')


def call_struct(name, args_struct, stream):
    """Entries that take one plain-old-data argument struct (include/phi3_b200.h) + the stream."""
    global launches
    L = lib()
    rc = getattr(L, name)(C.byref(args_struct), stream)
    launches += 1
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {L.p3_last_error().decode()}')


class GemmArgs(C.Structure):
    """p3_gemm_args (include/phi3_b200.h)"""
    _fields_ = [('X', _p), ('ldx', _l), ('W', _p), ('ldw', _l), ('bias', _p), ('out', _p), ('ldo', _l), ('resid', _p),
                ('row_map', _p), ('M', _l), ('N', C.c_int32), ('K', C.c_int32), ('epi', C.c_int32), ('impl', C.c_int32),
                ('ss_in', _p), ('n_ss_in', C.c_int32), ('eps', C.c_float), ('ss_out', _p), ('w_plan', _p),
                ('cosT', _p), ('sinT', _p), ('tab_bstride', _l),
                ('L', C.c_int32), ('n_heads', C.c_int32), ('n_kv', C.c_int32), ('hd', C.c_int32), ('past', C.c_int32),
                ('row_div', C.c_int32), ('write_cache', C.c_int32), ('bt_stride', C.c_int32),
                ('past_dev', _p), ('pool', _p), ('block_table', _p), ('splitk_ws', _p), ('splitk_ws_bytes', _l)]


class SkinnyArgs(C.Structure):
    """p3_skinny_args (include/phi3_b200.h)"""
    _fields_ = [('op', C.c_int32), ('X', _p), ('ldx', _l), ('norm_w', _p), ('eps', C.c_float), ('W', _p), ('Wq', _p), ('Wmeta', _p),
                ('out', _p), ('ldo', _l), ('resid', _p), ('M', C.c_int32), ('N', C.c_int32), ('K', C.c_int32), ('epi', C.c_int32),
                ('ss_in', _p), ('n_ss_in', C.c_int32), ('ss_out', _p), ('l2_prefetch', _p), ('l2_prefetch_bytes', _l),
                ('xg_gain', _p), ('xg_out', _p), ('ldxg', _l), ('rs_epi', C.c_int32),
                ('cosT', _p), ('sinT', _p), ('tab_bstride', _l),
                ('B', C.c_int32), ('L', C.c_int32), ('n_heads', C.c_int32), ('n_kv', C.c_int32), ('hd', C.c_int32), ('past', C.c_int32),
                ('row_div', C.c_int32), ('bt_stride', C.c_int32), ('write_cache', C.c_int32),
                ('past_dev', _p), ('pool', _p), ('block_table', _p), ('packed', C.c_int32)]


class WeightPlan:
    """caller-owned, 64-byte aligned blob holding the TMA descriptors of one weight matrix (p3_gemm_plan_weights)"""

    def __init__(self, w):
        self._raw = C.create_string_buffer(GEMM_WPLAN_BYTES + 64)
        self.addr = (C.addressof(self._raw) + 63) & ~63
        self.w = w                                             # keeps the tensor (and its address) alive
        rc = lib().p3_gemm_plan_weights(w.data_ptr(), w.stride(0), w.shape[0], w.shape[1], self.addr)
        if rc != 0:
            raise RuntimeError(f'p3_gemm_plan_weights failed ({rc}): {lib().p3_last_error().decode()}')
