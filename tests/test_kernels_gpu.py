"""-m gpu: every C-ABI kernel against a plain torch fp32 restatement of the same op
(floating-point kernels; tolerance stated per test) — the oracle-level parity tests are in
test_model_gpu.py."""
import math
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mods():
    import phi3_b200  # noqa
    from phi3_b200 import _lib
    return _lib


def bf(x):
    return x.to(torch.bfloat16)


def st():
    return torch.cuda.current_stream().cuda_stream


def test_rmsnorm(dev):
    L = _mods()
    torch.manual_seed(0)
    for T, H in [(5, 384), (33, 3072), (7, 8192)]:
        x = bf(torch.randn(T, H, device=dev) * 3)
        w = bf(1 + 0.1 * torch.randn(H, device=dev))
        y = torch.empty_like(x)
        L.call('p3_rmsnorm', x.data_ptr(), w.data_ptr(), y.data_ptr(), T, H, 1e-5, st())
        xf = x.float()
        ref = bf(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float())
        assert (y.float() - ref.float()).abs().max() <= 2 ** -7 * ref.float().abs().max()   # <= 1 bf16 ulp
        assert (y != ref).float().mean() < 0.01


def test_layernorm(dev):
    L = _mods()
    torch.manual_seed(0)
    T, H = 37, 1024
    x = torch.randn(T, H, device=dev) * 2 + 0.5
    w, b = bf(1 + 0.1 * torch.randn(H, device=dev)), bf(0.1 * torch.randn(H, device=dev))
    ref = torch.nn.functional.layer_norm(x, (H,), w.float(), b.float(), 1e-5)
    y = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
    L.call('p3_layernorm', x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), T, H, 1e-5, 0, st())
    assert (y.float() - ref).abs().max() < 3e-2
    y32 = torch.empty(T, H, device=dev)
    L.call('p3_layernorm', x.data_ptr(), w.data_ptr(), b.data_ptr(), y32.data_ptr(), T, H, 1e-5, 1, st())
    assert (y32 - ref).abs().max() < 1e-4


def test_embed_gather_negative_ids(dev):
    L = _mods()
    tab = bf(torch.randn(100, 64, device=dev))
    ids = torch.tensor([3, -1, 99, 0, -7, 250], dtype=torch.int32, device=dev)
    out = torch.empty(6, 64, device=dev, dtype=torch.bfloat16)
    L.call('p3_embed_gather', tab.data_ptr(), ids.data_ptr(), out.data_ptr(), 6, 64, 100, None, st())
    exp = tab[torch.tensor([3, 0, 99, 0, 0, 0], device=dev)]
    assert torch.equal(out, exp)


EPIS = ['none', 'bias', 'qgelu', 'gelu', 'resid', 'swiglu', 'f32', 'resid_f32', 'rowmap']


def _gemm_ref(x, w, epi, bias, resid):
    acc = x.float() @ w.float().T
    if bias is not None:
        acc = acc + bias.float()
    if epi == 'qgelu':
        return bf(acc * torch.sigmoid(1.702 * acc))
    if epi == 'gelu':
        return bf(torch.nn.functional.gelu(acc))
    if epi == 'resid':
        return bf(resid.float() + bf(acc).float())
    if epi == 'resid_f32':
        return resid + acc
    if epi == 'f32':
        return acc
    return bf(acc)


def _run_gemm(L, dev, M, N, K, epi, impl):
    torch.manual_seed(1)
    x = bf(torch.randn(M, K, device=dev))
    w = bf(torch.randn(N, K, device=dev) * K ** -0.5)
    bias = bf(torch.randn(N, device=dev)) if epi in ('bias', 'qgelu', 'gelu', 'resid_f32') else None
    code = dict(none=0, bias=0, qgelu=1, gelu=2, resid=3, swiglu=4, f32=5, resid_f32=6, rowmap=0)[epi]
    resid = row_map = None
    if epi == 'swiglu':
        from phi3_b200.model import interleave_gate_up
        out = torch.zeros(M, N // 2, device=dev, dtype=torch.bfloat16)
        wi = interleave_gate_up(w)
        L.call('p3_gemm', x.data_ptr(), K, wi.data_ptr(), K, None, out.data_ptr(), N // 2, None, None, M, N, K, 4, impl, st())
        acc = x.float() @ w.float().T
        g, u = bf(acc[:, :N // 2]).float(), bf(acc[:, N // 2:]).float()
        ref = bf(bf(torch.nn.functional.silu(g)).float() * u)
        return out, ref
    if epi in ('f32', 'resid_f32'):
        out = torch.zeros(M, N, device=dev)
        if epi == 'resid_f32':
            resid = torch.randn(M, N, device=dev)
            out.copy_(resid)
    else:
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        if epi == 'resid':
            resid = bf(torch.randn(M, N, device=dev))
            out.copy_(resid)
    ref = _gemm_ref(x, w, epi, bias, resid)
    nrows = M
    if epi == 'rowmap':
        perm = torch.randperm(M + 5, device=dev)[:M].to(torch.int32)
        row_map = perm
        out = torch.zeros(M + 5, N, device=dev, dtype=torch.bfloat16)
    L.call('p3_gemm', x.data_ptr(), K, w.data_ptr(), K, None if bias is None else bias.data_ptr(), out.data_ptr(), N,
           out.data_ptr() if resid is not None else None, None if row_map is None else row_map.data_ptr(), M, N, K, code,
           impl, st())
    if epi == 'rowmap':
        out = out[row_map.long()]
    return out, ref


def _check(out, ref, tol=2e-2):
    torch.cuda.synchronize()
    err = (out.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    assert math.isfinite(err) and err <= tol * scale, f'err {err} scale {scale}'


@pytest.mark.parametrize('epi', EPIS)
def test_gemm_mma_crosscheck(dev, epi):
    L = _mods()
    out, ref = _run_gemm(L, dev, 150, 512, 320, epi, impl=1)
    _check(out, ref)


@pytest.mark.parametrize('shape', [(128, 256, 64), (128, 256, 256), (300, 512, 384), (1000, 1024, 640),
                                   (257, 32064, 384), (2885, 3072, 1024), (64, 9216, 3072), (4096, 3072, 8192),
                                   (48, 3072, 3072), (100, 3072, 8192), (17, 1024, 4096)])
def test_gemm_tcgen05_shapes(dev, shape):
    L = _mods()
    M, N, K = shape
    out, ref = _run_gemm(L, dev, M, N, K, 'none', impl=0)
    _check(out, ref)


@pytest.mark.parametrize('epi', EPIS)
def test_gemm_tcgen05_epilogues(dev, epi):
    L = _mods()
    out, ref = _run_gemm(L, dev, 333, 1024, 512, epi, impl=0)
    _check(out, ref)
    out, ref = _run_gemm(L, dev, 2000, 8192 if epi == 'swiglu' else 3072, 384, epi, impl=0)   # BN=256 path
    _check(out, ref)


@pytest.mark.parametrize('M', [1, 4, 8, 11, 16])
@pytest.mark.parametrize('mode', ['plain', 'norm', 'resid', 'swiglu', 'f32'])
def test_gemm_skinny(dev, M, mode):
    L = _mods()
    torch.manual_seed(2)
    K = 3072 if mode != 'resid' else 8192
    N = {'plain': 9216, 'norm': 9216, 'resid': 3072, 'swiglu': 4096, 'f32': 32064}[mode]
    x = bf(torch.randn(M, K, device=dev))
    w = bf(torch.randn(N, K, device=dev) * K ** -0.5)
    nw = bf(1 + 0.1 * torch.randn(K, device=dev))
    xin = x
    if mode in ('norm', 'swiglu', 'f32'):
        xf = x.float()
        xin = bf(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * nw.float())
    acc = xin.float() @ w.float().T
    if mode == 'swiglu':
        from phi3_b200.model import interleave_gate_up
        out = torch.zeros(M, N // 2, device=dev, dtype=torch.bfloat16)
        L.call('p3_gemm_skinny', x.data_ptr(), K, nw.data_ptr(), 1e-5, interleave_gate_up(w).data_ptr(), out.data_ptr(),
               N // 2, None, M, N, K, 4, None, 0, None, None, 0, st())
        g, u = bf(acc[:, :N // 2]).float(), bf(acc[:, N // 2:]).float()
        ref = bf(bf(torch.nn.functional.silu(g)).float() * u)
    elif mode == 'f32':
        out = torch.zeros(M, N, device=dev)
        L.call('p3_gemm_skinny', x.data_ptr(), K, nw.data_ptr(), 1e-5, w.data_ptr(), out.data_ptr(), N, None, M, N, K, 5, None, 0, None, None, 0, st())
        ref = acc
    elif mode == 'resid':
        out = bf(torch.randn(M, N, device=dev))
        ref = bf(out.float() + bf(acc).float())
        L.call('p3_gemm_skinny', x.data_ptr(), K, None, 1e-5, w.data_ptr(), out.data_ptr(), N, out.data_ptr(), M, N, K, 3, None, 0, None, None, 0, st())
    else:
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        L.call('p3_gemm_skinny', x.data_ptr(), K, nw.data_ptr() if mode == 'norm' else None, 1e-5, w.data_ptr(),
               out.data_ptr(), N, None, M, N, K, 0, None, 0, None, None, 0, st())
        ref = bf(acc)
    _check(out, ref, tol=1e-2)


def _paged(kc, vc, dev):
    """dense [B,H,S,D] -> pool [pages,2,H,64,D], block_table"""
    B, H, S, D = kc.shape
    pps = (S + 63) // 64
    pool = torch.zeros(B * pps, 2, H, 64, D, device=dev, dtype=torch.bfloat16)
    bt = torch.arange(B * pps, dtype=torch.int32, device=dev).reshape(B, pps)
    bt = bt.flip(1).contiguous()               # non-trivial page order
    for b in range(B):
        for p in range(pps):
            n = min(64, S - p * 64)
            pool[bt[b, p], 0, :, :n] = kc[b, :, p * 64:p * 64 + n]
            pool[bt[b, p], 1, :, :n] = vc[b, :, p * 64:p * 64 + n]
    return pool, bt


def _attn_ref(q, k, v, scale, causal, past, kv_start):
    """q [B,H,L,D], k/v [B,H,S,D] (S = past+L)."""
    B, H, Lq, D = q.shape
    S = k.shape[2]
    s = (q.float() * scale) @ k.float().transpose(-1, -2)
    qi = past + torch.arange(Lq, device=q.device)[:, None]
    kj = torch.arange(S, device=q.device)[None, :]
    allow = torch.ones(Lq, S, dtype=torch.bool, device=q.device)
    if causal:
        allow = kj <= qi
    allow = allow[None, None] & (kj[None, None] >= kv_start[:, None, None, None])
    s = s.masked_fill(~allow, float('-inf'))
    dead = (~allow).all(-1, keepdim=True)
    p = torch.softmax(s.masked_fill(dead, 0), -1).masked_fill(dead, 0)
    return p @ v.float()


@pytest.mark.parametrize('cfg', [(2, 4, 96, 200, 0, True), (1, 16, 64, 577, 0, False), (3, 2, 96, 64, 0, True),
                                 (2, 3, 96, 130, 100, True), (1, 2, 96, 1100, 0, True)])
def test_attention_prefill(dev, cfg):
    L = _mods()
    B, H, D, Lq, past, causal = cfg
    torch.manual_seed(3)
    qkv = bf(torch.randn(B * Lq, 3 * H * D, device=dev))
    kc, vc = bf(torch.randn(B, H, past, D, device=dev)), bf(torch.randn(B, H, past, D, device=dev))
    kv_start = torch.tensor([0, 17, 70][:B], dtype=torch.int32, device=dev) if causal else torch.zeros(B, dtype=torch.int32, device=dev)
    out = torch.zeros(B * Lq, H * D, device=dev, dtype=torch.bfloat16)
    pool, bt = _paged(kc, vc, dev) if past else (None, None)
    p = qkv.data_ptr()
    L.call('p3_attention_prefill', p, p + H * D * 2, p + 2 * H * D * 2, 3 * H * D, 3 * H * D, 3 * H * D, out.data_ptr(),
           H * D, B, Lq, H, H, D, D ** -0.5, int(causal), past, kv_start.data_ptr(),
           None if pool is None else pool.data_ptr(), None if bt is None else bt.data_ptr(), 0 if bt is None else bt.stride(0), 1, st())
    x = qkv.view(B, Lq, 3, H, D).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], torch.cat([kc, x[1]], 2), torch.cat([vc, x[2]], 2)
    ref = _attn_ref(q, k, v, D ** -0.5, causal, past, kv_start.long()).transpose(1, 2).reshape(B * Lq, H * D)
    _check(out, ref, tol=2e-2)


def _prefill_call(L, qkv, kc, vc, kv_start, B, H, D, Lq, past, causal, dev):
    out = torch.full((B * Lq, H * D), float('nan'), device=dev, dtype=torch.bfloat16)
    pool, bt = _paged(kc, vc, dev) if past else (None, None)
    p = qkv.data_ptr()
    L.call('p3_attention_prefill', p, p + H * D * 2, p + 2 * H * D * 2, 3 * H * D, 3 * H * D, 3 * H * D, out.data_ptr(),
           H * D, B, Lq, H, H, D, D ** -0.5, int(causal), past, kv_start.data_ptr(),
           None if pool is None else pool.data_ptr(), None if bt is None else bt.data_ptr(), 0 if bt is None else bt.stride(0), 1, st())
    torch.cuda.synchronize()
    return out


# (B, H, D, L, past, causal, kv_start, key-magnitude ramp per 128-key tile)
_TC_CFGS = [(1, 1, 64, 128, 0, False, [0], 1.0), (1, 2, 96, 256, 0, True, [0], 1.0), (2, 3, 96, 300, 256, True, [0, 200], 1.0),
            (1, 2, 96, 640, 128, True, [5], 1.0), (3, 2, 64, 577, 0, False, [0, 0, 0], 1.0), (1, 4, 96, 2048, 0, True, [0], 1.0),
            (2, 2, 96, 1100, 0, True, [0, 300], 1.0), (1, 2, 96, 1024, 0, True, [0], 3.0), (1, 2, 96, 768, 0, False, [0], 40.0),
            (2, 2, 64, 700, 0, False, [0, 0], 6.0), (1, 2, 96, 1000, 384, True, [130], 1.0),
            (2, 2, 96, 333, 192, True, [10, 70], 1.0), (1, 3, 96, 64, 64, True, [0], 1.0)]


@pytest.mark.parametrize('cfg', _TC_CFGS)
def test_attention_prefill_tcgen05(dev, cfg):
    """tcgen05 flash attention (attention_tc.cu) vs the fp32 restatement and vs the mma.sync kernel on the same inputs:
    paged past (past % 64 == 0), left padding, ragged last tiles, ViT shape, and score ramps that drive the
    lazy-rescale (> 2^8) and redo (> 2^64) paths of the streaming softmax."""
    import os
    L = _mods()
    B, H, D, Lq, past, causal, kv, ramp = cfg
    torch.manual_seed(5)
    qkv = bf(torch.randn(B * Lq, 3 * H * D, device=dev))
    if ramp != 1.0:
        r = torch.tensor([ramp ** (i // 128) for i in range(Lq)], device=dev)
        qkv.view(B, Lq, 3, H, D)[:, :, 1].mul_(r[None, :, None, None].to(torch.bfloat16))
    kc, vc = bf(torch.randn(B, H, past, D, device=dev)), bf(torch.randn(B, H, past, D, device=dev))
    kv_start = torch.tensor(kv, dtype=torch.int32, device=dev)
    x = qkv.view(B, Lq, 3, H, D).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], torch.cat([kc, x[1]], 2), torch.cat([vc, x[2]], 2)
    ref = _attn_ref(q, k, v, D ** -0.5, causal, past, kv_start.long()).transpose(1, 2).reshape(B * Lq, H * D)
    old = os.environ.get('P3_ATTN_TC')
    try:
        os.environ['P3_ATTN_TC'] = '1'
        out_tc = _prefill_call(L, qkv, kc, vc, kv_start, B, H, D, Lq, past, causal, dev)
        os.environ['P3_ATTN_TC'] = '0'
        out_mma = _prefill_call(L, qkv, kc, vc, kv_start, B, H, D, Lq, past, causal, dev)
    finally:
        if old is None:
            os.environ.pop('P3_ATTN_TC', None)
        else:
            os.environ['P3_ATTN_TC'] = old
    assert not torch.isnan(out_tc.float()).any()
    _check(out_tc, ref, tol=1e-2)
    _check(out_mma, ref, tol=1e-2)
    _check(out_tc, out_mma.float(), tol=1e-2)


@pytest.mark.parametrize('cfg', [(2, 4, 96, 1, 300, 1, 1), (2, 4, 96, 1, 300, 3, 1), (3, 2, 96, 5, 1000, 4, 1),
                                 (4, 2, 96, 6, 257, 2, 2), (1, 32, 96, 1, 2100, 5, 1), (2, 2, 96, 16, 64, 1, 1),
                                 (2, 2, 96, 3, 0, 1, 1), (2, 4, 64, 2, 130, 2, 1)])
def test_attention_decode(dev, cfg):
    L = _mods()
    B, H, D, Lq, past, n_splits, n_beam = cfg
    torch.manual_seed(4)
    nseq = B // n_beam
    qkv = bf(torch.randn(B * Lq, 3 * H * D, device=dev))
    kc, vc = bf(torch.randn(nseq, H, past, D, device=dev)), bf(torch.randn(nseq, H, past, D, device=dev))
    kv_start = torch.tensor([0, 9, 70, 3][:nseq], dtype=torch.int32, device=dev).clamp(max=max(past - 1, 0))
    pool, bt = _paged(kc, vc, dev) if past else (torch.zeros(1, 2, H, 64, D, device=dev, dtype=torch.bfloat16),
                                                 torch.zeros(nseq, 1, dtype=torch.int32, device=dev))
    out = torch.zeros(B * Lq, H * D, device=dev, dtype=torch.bfloat16)
    nb = L.lib().p3_attention_decode_workspace(B, Lq, H, D, n_splits)
    ws = torch.zeros(nb // 4, device=dev)
    p = qkv.data_ptr()
    L.call('p3_attention_decode', p, p + H * D * 2, p + 2 * H * D * 2, 3 * H * D, 3 * H * D, 3 * H * D, out.data_ptr(),
           H * D, B, Lq, H, H, D, D ** -0.5, past, kv_start.data_ptr(), pool.data_ptr(), bt.data_ptr(), bt.stride(0),
           n_beam, n_splits, ws.data_ptr(), None, None, 0, st())
    x = qkv.view(B, Lq, 3, H, D).permute(2, 0, 3, 1, 4)
    kr, vr = kc.repeat_interleave(n_beam, 0), vc.repeat_interleave(n_beam, 0)
    q, k, v = x[0], torch.cat([kr, x[1]], 2), torch.cat([vr, x[2]], 2)
    ref = _attn_ref(q, k, v, D ** -0.5, True, past, kv_start.long().repeat_interleave(n_beam)).transpose(1, 2).reshape(B * Lq, H * D)
    _check(out, ref, tol=2e-2)


def test_rope_kvwrite(dev):
    L = _mods()
    torch.manual_seed(5)
    B, Lq, H, D, past, S = 2, 5, 4, 96, 70, 160
    qkv = bf(torch.randn(B * Lq, 3 * H * D, device=dev))
    orig = qkv.clone()
    ang = torch.rand(B, S, D // 2, device=dev) * 6
    cos, sin = (torch.cos(ang) * 1.19).contiguous(), (torch.sin(ang) * 1.19).contiguous()
    pps = (S + 63) // 64
    pool = torch.zeros(B * pps, 2, H, 64, D, device=dev, dtype=torch.bfloat16)
    bt = torch.arange(B * pps, dtype=torch.int32, device=dev).reshape(B, pps)
    L.call('p3_rope_kvwrite', qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), S * (D // 2), B, Lq, H, H, D, past, 1,
           pool.data_ptr(), bt.data_ptr(), pps, 1, None, st())
    x = orig.view(B, Lq, 3, H, D).float()
    c = torch.cat([cos, cos], -1)[:, past:past + Lq, None, :]
    s_ = torch.cat([sin, sin], -1)[:, past:past + Lq, None, :]
    rot = lambda t: t * c + torch.cat([-t[..., D // 2:], t[..., :D // 2]], -1) * s_
    qr, kr = bf(rot(x[:, :, 0])), bf(rot(x[:, :, 1]))
    got = qkv.view(B, Lq, 3, H, D)
    assert (got[:, :, 0].float() - qr.float()).abs().max() <= 2 ** -6 * qr.float().abs().max()
    assert (got[:, :, 1].float() - kr.float()).abs().max() <= 2 ** -6 * kr.float().abs().max()
    assert torch.equal(got[:, :, 2], orig.view(B, Lq, 3, H, D)[:, :, 2])
    for b in range(B):
        for i in range(Lq):
            pos = past + i
            pg = bt[b, pos // 64]
            assert torch.equal(pool[pg, 0, :, pos % 64], got[b, i, 1])
            assert torch.equal(pool[pg, 1, :, pos % 64], got[b, i, 2])


def test_row_stats(dev):
    L = _mods()
    torch.manual_seed(6)
    R, V = 7, 32064
    lg = torch.randn(R, V, device=dev) * 3
    lg[2, 100] = lg[2, 50] = 20.0                      # tie -> first index
    from phi3_b200.api import _row_stats

    class M:
        pass
    g = torch.randint(0, V, (R, 5))
    out = _row_stats(M(), lg, n_top=4, gather=g)
    lp = torch.log_softmax(lg, -1)
    assert torch.equal(out['argmax'].long().cpu(), lg.argmax(-1).cpu())
    assert out['argmax'][2].item() == 50
    assert (out['lse'] - torch.logsumexp(lg, -1)).abs().max() < 1e-3
    tv, ti = lp.topk(4, -1)
    assert torch.equal(out['top_ids'].long().sort(-1).values, ti.sort(-1).values)
    assert (out['top_lp'] - tv).abs().max() < 1e-3
    assert (out['gather_lp'] - lp.gather(1, g.to(dev))).abs().max() < 1e-3


def test_kv_quant_roundtrip_matches_oracle(dev):
    L = _mods()
    from oracle.phi3_oracle import quantize_q4g32, dequantize_q4g32
    torch.manual_seed(7)
    nseq, H, D, S = 2, 3, 96, 150
    kc, vc = bf(torch.randn(nseq, H, S, D, device=dev) * 2), bf(torch.randn(nseq, H, S, D, device=dev))
    pool, bt = _paged(kc, vc, dev)
    qc = torch.zeros(pool.shape[0], 2, H, 64, D // 2, dtype=torch.uint8, device=dev)
    qm = torch.zeros(pool.shape[0], 2, H, 64, D // 32, 2, dtype=torch.bfloat16, device=dev)
    L.call('p3_kv_quantize_q4g32', pool.data_ptr(), qc.data_ptr(), qm.data_ptr(), bt.data_ptr(), bt.stride(0), nseq, S, H, D, st())
    torch.cuda.synchronize()
    for src, kv in ((kc, 0), (vc, 1)):
        flat = src.float().cpu().reshape(nseq * H, -1)
        deq = dequantize_q4g32(*quantize_q4g32(flat, 'b200'), (nseq, H, S, D)).to(torch.bfloat16)
        for b in range(nseq):
            for p in range((S + 63) // 64):
                n = min(64, S - p * 64)
                got = pool[bt[b, p], kv, :, :n].cpu()
                assert torch.equal(got, deq[b, :, p * 64:p * 64 + n]), (kv, b, p)


def _dequant_pool(qc, qm):
    """codes [pages,2,H,64,D/2] u8 (low nibble = even dim) + meta [pages,2,H,64,D/32,2] bf16 -> bf16(code*scale + bias)."""
    lo, hi = (qc & 15).float(), (qc >> 4).float()
    codes = torch.stack([lo, hi], -1).reshape(*qc.shape[:-1], -1)                      # [..., D]
    sc = qm[..., 0].float().repeat_interleave(32, -1)
    bi = qm[..., 1].float().repeat_interleave(32, -1)
    return (codes * sc + bi).to(torch.bfloat16)


# (B, H, L, past, kv_start, n_splits): plain decode, beam-sized L (rows 8..15 live), left padding into the quantised
# pages, a partial bf16 tail page, only-quantised / only-bf16 histories, split merge
_Q4_CFGS = [(2, 3, 1, 200, [0, 0], 1), (2, 2, 5, 330, [0, 37], 2), (1, 2, 16, 130, [0], 1), (3, 2, 9, 197, [0, 70, 3], 1),
            (2, 2, 1, 256, [0, 130], 4), (1, 2, 3, 50, [7], 1), (2, 2, 8, 640, [0, 64], 3), (1, 1, 12, 64, [0], 1)]


@pytest.mark.parametrize('cfg', _Q4_CFGS)
def test_decode_attention_q4_matches_dequantised_reference(dev, cfg):
    """Quantised-cache decode attention (codes through the MMA, affine map on group sums) vs fp32 attention over the
    dequantised cache bf16(code*scale+bias) — the oracle's order of operations (phi:536-539)."""
    L = _mods()
    B, H, Lq, past, kv, ns = cfg
    D = 96
    torch.manual_seed(21)
    qkv = bf(torch.randn(B * Lq, 3 * H * D, device=dev))
    kc = bf(torch.randn(B, H, past, D, device=dev) * 1.5 + 0.3)
    vc = bf(torch.randn(B, H, past, D, device=dev) + torch.linspace(-2, 2, D, device=dev))
    pool, bt = _paged(kc, vc, dev)
    n_quant = (past // 64) * 64
    qc = torch.zeros(pool.shape[0], 2, H, 64, D // 2, dtype=torch.uint8, device=dev)
    qm = torch.zeros(pool.shape[0], 2, H, 64, D // 32, 2, dtype=torch.bfloat16, device=dev)
    if n_quant:
        L.call('p3_kv_quantize_q4g32', pool.data_ptr(), qc.data_ptr(), qm.data_ptr(), bt.data_ptr(), bt.stride(0), B, n_quant, H, D, st())
    kv_start = torch.tensor(kv, dtype=torch.int32, device=dev)
    out = torch.full((B * Lq, H * D), float('nan'), device=dev, dtype=torch.bfloat16)
    ws = torch.zeros(max(L.lib().p3_attention_decode_workspace(B, Lq, H, D, ns) // 4, 1), device=dev)
    p = qkv.data_ptr()
    L.call('p3_attention_decode_q4', p, p + H * D * 2, p + 2 * H * D * 2, 3 * H * D, 3 * H * D, 3 * H * D, out.data_ptr(), H * D,
           B, Lq, H, H, D, D ** -0.5, past, n_quant, kv_start.data_ptr(), pool.data_ptr(), qc.data_ptr(), qm.data_ptr(),
           bt.data_ptr(), bt.stride(0), 1, ns, ws.data_ptr(), None, None, 0, st())
    torch.cuda.synchronize()
    # what the kernel is specified to see: dequantised pages for [0, n_quant), bf16 pages after
    deq = _dequant_pool(qc, qm)
    kd, vd = kc.clone(), vc.clone()
    for b in range(B):
        for pg in range(n_quant // 64):
            kd[b, :, pg * 64:(pg + 1) * 64] = deq[bt[b, pg], 0]
            vd[b, :, pg * 64:(pg + 1) * 64] = deq[bt[b, pg], 1]
    x = qkv.view(B, Lq, 3, H, D).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], torch.cat([kd, x[1]], 2), torch.cat([vd, x[2]], 2)
    ref = _attn_ref(q, k, v, D ** -0.5, True, past, kv_start.long()).transpose(1, 2).reshape(B * Lq, H * D)
    assert not torch.isnan(out.float()).any()
    _check(out, ref, tol=1e-2)
    # and the quantisation itself is visible: the un-quantised cache gives a different answer when pages were quantised
    if n_quant:
        ref_bf = _attn_ref(q, torch.cat([kc, x[1]], 2), torch.cat([vc, x[2]], 2), D ** -0.5, True, past, kv_start.long())
        assert (ref_bf.transpose(1, 2).reshape(B * Lq, H * D) - ref).abs().max() > 1e-3


def test_skinny_ss_partials_feed_fused_rmsnorm(dev):
    """RESIDUAL epilogue emits per-CTA sum-of-squares partials; a following norm-fused skinny GEMM fed
    with them must equal the one that recomputes the statistic from X."""
    L = _mods()
    torch.manual_seed(8)
    M, H = 7, 3072
    act = bf(torch.randn(M, H, device=dev))
    wo = bf(torch.randn(H, H, device=dev) * H ** -0.5)
    h = bf(torch.randn(M, H, device=dev))
    ss = torch.zeros((H // 16, 16), device=dev)
    L.call('p3_gemm_skinny', act.data_ptr(), H, None, 1e-5, wo.data_ptr(), h.data_ptr(), H, h.data_ptr(), M, H, H, 3,
           None, 0, ss.data_ptr(), None, 0, st())
    torch.cuda.synchronize()
    ref_ss = h.float().pow(2).sum(-1)
    got_ss = ss.sum(0)[:M]
    assert (got_ss - ref_ss).abs().max() <= 1e-3 * ref_ss.max()
    assert (ss[:, M:] == 0).all()
    nw = bf(1 + 0.1 * torch.randn(H, device=dev))
    w2 = bf(torch.randn(1024, H, device=dev) * H ** -0.5)
    o1 = torch.zeros(M, 1024, device=dev, dtype=torch.bfloat16)
    o2 = torch.zeros_like(o1)
    L.call('p3_gemm_skinny', h.data_ptr(), H, nw.data_ptr(), 1e-5, w2.data_ptr(), o1.data_ptr(), 1024, None, M, 1024, H, 0,
           ss.data_ptr(), H // 16, None, None, 0, st())
    L.call('p3_gemm_skinny', h.data_ptr(), H, nw.data_ptr(), 1e-5, w2.data_ptr(), o2.data_ptr(), 1024, None, M, 1024, H, 0,
           None, 0, None, None, 0, st())
    torch.cuda.synchronize()
    assert (o1.float() - o2.float()).abs().max() <= 2 ** -7 * o2.float().abs().max()
    assert (o1 != o2).float().mean() < 0.02
    # embed_gather's ss_out
    tab = bf(torch.randn(50, H, device=dev))
    ids = torch.tensor([3, 7, 49], dtype=torch.int32, device=dev)
    out = torch.empty(3, H, device=dev, dtype=torch.bfloat16)
    sse = torch.zeros(16, device=dev)
    L.call('p3_embed_gather', tab.data_ptr(), ids.data_ptr(), out.data_ptr(), 3, H, 50, sse.data_ptr(), st())
    torch.cuda.synchronize()
    assert (sse[:3] - tab[ids.long()].float().pow(2).sum(-1)).abs().max() < 1e-2


@pytest.mark.parametrize('w4', [False, True])
@pytest.mark.parametrize('M', [3, 8, 13])
def test_skinny_split_rmsnorm_matches_fused(dev, M, w4):
    """p3_gemm_skinny_x: producer writes bf16(h * gain) next to h (+ sum-of-squares partials), the consumer applies
    rsqrt(mean(h^2) + eps) to its accumulators == the RMSNorm-fused skinny GEMM on h (up to the place of one bf16 rounding)."""
    import ctypes as C
    from phi3_b200 import quant
    L = _mods()
    torch.manual_seed(31 + M)
    H, N2 = 3072, 1024
    act = bf(torch.randn(M, H, device=dev))
    wo = bf(torch.randn(H, H, device=dev) * H ** -0.5)
    h0 = bf(torch.randn(M, H, device=dev))
    gain = bf(1 + 0.2 * torch.randn(H, device=dev))
    w2 = bf(torch.randn(N2, H, device=dev) * H ** -0.5)
    q2 = quant.W4(w2, pack=True) if w4 else None

    def args(**kw):
        a = L.SkinnyArgs()
        a.eps = 1e-5
        for k, v in kw.items():
            setattr(a, k, v)
        return a
    # producer: o_proj-like residual launch, also emitting hg and the statistics
    h = h0.clone()
    hg = torch.zeros_like(h)
    ss = torch.zeros((H // 16, 16), device=dev)
    L.call_struct('p3_gemm_skinny_x', args(op=0, X=act.data_ptr(), ldx=H, W=wo.data_ptr(), out=h.data_ptr(), ldo=H, resid=h.data_ptr(),
                                           M=M, N=H, K=H, epi=3, ss_out=ss.data_ptr(), xg_gain=gain.data_ptr(), xg_out=hg.data_ptr(), ldxg=H), st())
    torch.cuda.synchronize()
    h_ref = (h0.float() + bf(act.float() @ wo.float().T).float()).to(torch.bfloat16)
    assert (h.float() - h_ref.float()).abs().max() <= 2 ** -7 * h_ref.float().abs().max()
    assert torch.equal(hg, (h.float() * gain.float()).to(torch.bfloat16))
    # consumer: rs_epi on hg vs the fused-norm kernel on h
    o1 = torch.zeros(M, N2, device=dev, dtype=torch.bfloat16)
    o2 = torch.zeros_like(o1)
    wkw = dict(Wq=q2.codes.data_ptr(), Wmeta=q2.meta.data_ptr()) if w4 else dict(W=w2.data_ptr())
    L.call_struct('p3_gemm_skinny_x', args(op=0, X=hg.data_ptr(), ldx=H, out=o1.data_ptr(), ldo=N2, M=M, N=N2, K=H, epi=0,
                                           ss_in=ss.data_ptr(), n_ss_in=H // 16, rs_epi=1, **wkw), st())
    if w4:
        L.call('p3_gemm_skinny_w4', h.data_ptr(), H, gain.data_ptr(), 1e-5, q2.codes.data_ptr(), q2.meta.data_ptr(), o2.data_ptr(), N2, None,
               M, N2, H, 0, ss.data_ptr(), H // 16, None, None, 0, st())
    else:
        L.call('p3_gemm_skinny', h.data_ptr(), H, gain.data_ptr(), 1e-5, w2.data_ptr(), o2.data_ptr(), N2, None, M, N2, H, 0,
               ss.data_ptr(), H // 16, None, None, 0, st())
    torch.cuda.synchronize()
    wd = q2.deq.float() if w4 else w2.float()
    hn = h.float() * torch.rsqrt(h.float().pow(2).mean(-1, keepdim=True) + 1e-5) * gain.float()
    ref = hn @ wd.T
    assert (o1.float() - ref).abs().max() <= 1e-2 * ref.abs().max()
    assert (o1.float() - o2.float()).abs().max() <= 1e-2 * ref.abs().max()


@pytest.mark.parametrize('K', [3072, 8192])
@pytest.mark.parametrize('M', [1, 8, 13])
def test_skinny_packed_tile_order_is_bit_identical(dev, M, K):
    """p3_gemm_skinny_x with W in the stream order of the 16-row tiles (model.pack_rows16) == the row-major launch, bit for bit
    (same loads per lane, same arithmetic order): output, gain-scaled copy and the sum-of-squares partials."""
    from phi3_b200.model import pack_rows16
    L = _mods()
    torch.manual_seed(K + M)
    N = 3072
    x = bf(torch.randn(M, K, device=dev))
    w = bf(torch.randn(N, K, device=dev) * K ** -0.5)
    wp = pack_rows16(w)
    assert wp.shape == w.shape and not torch.equal(wp, w)
    # element (n, k) sits at tile n // 16, chunk k // 64, row half (n % 16) // 8, k half (k % 64) // 32, lane 4 * (n % 8) + (k % 32) // 8
    n, k = 37, 1000
    flat = (((n // 16) * (K // 64) + k // 64) * 4 + ((n % 16) // 8) * 2 + (k % 64) // 32) * 256 + (4 * (n % 8) + (k % 32) // 8) * 8 + k % 8
    assert wp.view(-1)[flat] == w[n, k]
    h0 = bf(torch.randn(M, N, device=dev))
    gain = bf(1 + 0.2 * torch.randn(N, device=dev))
    outs = []
    for W, packed in ((w, 0), (wp, 1)):
        h, hg, ss = h0.clone(), torch.zeros(M, N, device=dev, dtype=torch.bfloat16), torch.zeros((N // 16, 16), device=dev)
        a = L.SkinnyArgs()
        a.op, a.X, a.ldx, a.W, a.out, a.ldo, a.resid, a.M, a.N, a.K, a.epi = 0, x.data_ptr(), K, W.data_ptr(), h.data_ptr(), N, h.data_ptr(), M, N, K, 3
        a.ss_out, a.xg_gain, a.xg_out, a.ldxg, a.packed, a.eps = ss.data_ptr(), gain.data_ptr(), hg.data_ptr(), N, packed, 1e-5
        L.call_struct('p3_gemm_skinny_x', a, st())
        torch.cuda.synchronize()
        outs.append((h, hg, ss))
    for u, v in zip(*outs):
        assert torch.equal(u, v)
    ref = (h0.float() + bf(x.float() @ w.float().T).float()).to(torch.bfloat16)
    assert (outs[1][0].float() - ref.float()).abs().max() <= 2 ** -6 * ref.float().abs().max()
    # the packed layout is refused where the kernel would not read 16-row tiles
    a.epi = 4
    with pytest.raises(RuntimeError):
        L.call_struct('p3_gemm_skinny_x', a, st())


@pytest.mark.parametrize('M', [(2, 1), (8, 1), (2, 5), (12, 1)])
def test_fused_qkv_rope_matches_unfused(dev, M):
    """p3_gemm_skinny_qkv_rope == p3_gemm_skinny (norm fused) followed by p3_rope_kvwrite."""
    L = _mods()
    B, Lq = M
    T, H, nh, D, past, S = B * Lq, 3072, 32, 96, 70, 160
    torch.manual_seed(9)
    x = bf(torch.randn(T, H, device=dev))
    nw = bf(1 + 0.1 * torch.randn(H, device=dev))
    w = bf(torch.randn(3 * nh * D, H, device=dev) * H ** -0.5)
    ang = torch.rand(B, S, D // 2, device=dev) * 6
    cos, sin = (torch.cos(ang) * 1.19).contiguous(), (torch.sin(ang) * 1.19).contiguous()
    pps = (S + 63) // 64
    bt = torch.arange(B * pps, dtype=torch.int32, device=dev).reshape(B, pps)
    pool1 = torch.zeros(B * pps, 2, nh, 64, D, device=dev, dtype=torch.bfloat16)
    pool2 = torch.zeros_like(pool1)
    q1 = torch.zeros(T, 3 * nh * D, device=dev, dtype=torch.bfloat16)
    q2 = torch.zeros_like(q1)
    L.call('p3_gemm_skinny', x.data_ptr(), H, nw.data_ptr(), 1e-5, w.data_ptr(), q1.data_ptr(), 3 * nh * D, None, T, 3 * nh * D, H, 0,
           None, 0, None, None, 0, st())
    L.call('p3_rope_kvwrite', q1.data_ptr(), cos.data_ptr(), sin.data_ptr(), S * (D // 2), B, Lq, nh, nh, D, past, 1,
           pool1.data_ptr(), bt.data_ptr(), pps, 1, None, st())
    L.call('p3_gemm_skinny_qkv_rope', x.data_ptr(), H, nw.data_ptr(), 1e-5, w.data_ptr(), q2.data_ptr(), None, 0, cos.data_ptr(),
           sin.data_ptr(), S * (D // 2), B, Lq, nh, nh, D, H, past, None, 1, pool2.data_ptr(), bt.data_ptr(), pps, 1, None, 0, st())
    torch.cuda.synchronize()
    tol = 2 ** -7 * q1.float().abs().max()
    assert (q1.float() - q2.float()).abs().max() <= tol
    assert (q1 != q2).float().mean() < 0.01
    assert (pool1.float() - pool2.float()).abs().max() <= tol
    assert (pool2 != 0).any()


@pytest.mark.parametrize('M', [1, 5, 8, 13, 16])
@pytest.mark.parametrize('mode', ['plain', 'norm', 'resid', 'swiglu', 'f32'])
def test_gemm_skinny_w4(dev, M, mode):
    """quantize_model decode GEMM (4-bit g64 codes, affine map applied to per-group tensor-core sums) vs
    x @ dequantise(quantise(W)).T computed by torch in fp32."""
    L = _mods()
    from phi3_b200 import quant
    from phi3_b200.model import interleave_gate_up
    torch.manual_seed(12)
    K = 3072 if mode != 'resid' else 8192
    N = {'plain': 9216, 'norm': 9216, 'resid': 3072, 'swiglu': 4096, 'f32': 32064}[mode]
    x = bf(torch.randn(M, K, device=dev))
    w = bf(torch.randn(N, K, device=dev) * K ** -0.5)
    nw = bf(1 + 0.1 * torch.randn(K, device=dev))
    q = quant.W4(w, row_perm=interleave_gate_up if mode == 'swiglu' else None)
    codes, scale, bias = quant.quantize_w4g64(w)
    wq = (codes.reshape(N, K // 64, 64).float() * scale.float()[..., None] + bias.float()[..., None]).reshape(N, K)   # unrounded image
    xin = x
    if mode in ('norm', 'swiglu', 'f32'):
        xf = x.float()
        xin = bf(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * nw.float())
    acc = xin.float() @ wq.T
    c, m = q.codes.data_ptr(), q.meta.data_ptr()
    if mode == 'swiglu':
        out = torch.zeros(M, N // 2, device=dev, dtype=torch.bfloat16)
        L.call('p3_gemm_skinny_w4', x.data_ptr(), K, nw.data_ptr(), 1e-5, c, m, out.data_ptr(), N // 2, None, M, N, K, 4, None, 0,
               None, None, 0, st())
        g, u = bf(acc[:, :N // 2]).float(), bf(acc[:, N // 2:]).float()
        ref = bf(bf(torch.nn.functional.silu(g)).float() * u)
    elif mode == 'f32':
        out = torch.zeros(M, N, device=dev)
        L.call('p3_gemm_skinny_w4', x.data_ptr(), K, nw.data_ptr(), 1e-5, c, m, out.data_ptr(), N, None, M, N, K, 5, None, 0, None, None, 0, st())
        ref = acc
    elif mode == 'resid':
        out = bf(torch.randn(M, N, device=dev))
        ref = bf(out.float() + bf(acc).float())
        L.call('p3_gemm_skinny_w4', x.data_ptr(), K, None, 1e-5, c, m, out.data_ptr(), N, out.data_ptr(), M, N, K, 3, None, 0, None, None, 0, st())
    else:
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        L.call('p3_gemm_skinny_w4', x.data_ptr(), K, nw.data_ptr() if mode == 'norm' else None, 1e-5, c, m,
               out.data_ptr(), N, None, M, N, K, 0, None, 0, None, None, 0, st())
        ref = bf(acc)
    _check(out, ref, tol=1e-2)
    # and it is the same op as the bf16 kernel over the dequantised image (what the prefill GEMMs use)
    if mode == 'plain':
        out2 = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        L.call('p3_gemm_skinny', x.data_ptr(), K, None, 1e-5, q.deq.data_ptr(), out2.data_ptr(), N, None, M, N, K, 0, None, 0,
               None, None, 0, st())
        _check(out, out2.float(), tol=1e-2)


@pytest.mark.parametrize('shape', [(3, 1000, 256), (16, 40, 128), (7, 3080, 1024), (9, 384, 384)])
def test_gemm_skinny_w4_ragged_shapes(dev, shape):
    """N that is not a multiple of the 16/32-row CTA tile, the smallest K (one 128-wide super-chunk), odd token counts"""
    L = _mods()
    from phi3_b200 import quant
    M, N, K = shape
    torch.manual_seed(21)
    x = bf(torch.randn(M, K, device=dev))
    w = bf(torch.randn(N, K, device=dev) * K ** -0.5)
    q = quant.W4(w)
    out = torch.full((M, N), float('nan'), device=dev, dtype=torch.bfloat16)
    L.call('p3_gemm_skinny_w4', x.data_ptr(), K, None, 1e-5, q.codes.data_ptr(), q.meta.data_ptr(), out.data_ptr(), N, None,
           M, N, K, 0, None, 0, None, None, 0, st())
    _check(out, bf(x.float() @ q.deq.float().T), tol=1e-2)
    with pytest.raises(RuntimeError):                       # K must be a multiple of 128 for the 4-bit stream
        L.call('p3_gemm_skinny_w4', x.data_ptr(), K, None, 1e-5, q.codes.data_ptr(), q.meta.data_ptr(), out.data_ptr(), N, None,
               M, N, 192, 0, None, 0, None, None, 0, st())


@pytest.mark.parametrize('M', [(1, 1), (8, 1), (2, 5)])
def test_fused_qkv_rope_w4_matches_bf16_kernel_on_dequantised_weights(dev, M):
    L = _mods()
    from phi3_b200 import quant
    B, Lq = M
    T, H, nh, D, past, S = B * Lq, 3072, 32, 96, 70, 160
    torch.manual_seed(19)
    x = bf(torch.randn(T, H, device=dev))
    nw = bf(1 + 0.1 * torch.randn(H, device=dev))
    q = quant.W4(bf(torch.randn(3 * nh * D, H, device=dev) * H ** -0.5))
    ang = torch.rand(B, S, D // 2, device=dev) * 6
    cos, sin = (torch.cos(ang) * 1.19).contiguous(), (torch.sin(ang) * 1.19).contiguous()
    pps = (S + 63) // 64
    bt = torch.arange(B * pps, dtype=torch.int32, device=dev).reshape(B, pps)
    pool1 = torch.zeros(B * pps, 2, nh, 64, D, device=dev, dtype=torch.bfloat16)
    pool2 = torch.zeros_like(pool1)
    q1 = torch.zeros(T, 3 * nh * D, device=dev, dtype=torch.bfloat16)
    q2 = torch.zeros_like(q1)
    L.call('p3_gemm_skinny_qkv_rope', x.data_ptr(), H, nw.data_ptr(), 1e-5, q.deq.data_ptr(), q1.data_ptr(), None, 0, cos.data_ptr(),
           sin.data_ptr(), S * (D // 2), B, Lq, nh, nh, D, H, past, None, 1, pool1.data_ptr(), bt.data_ptr(), pps, 1, None, 0, st())
    L.call('p3_gemm_skinny_qkv_rope_w4', x.data_ptr(), H, nw.data_ptr(), 1e-5, q.codes.data_ptr(), q.meta.data_ptr(), q2.data_ptr(),
           None, 0, cos.data_ptr(), sin.data_ptr(), S * (D // 2), B, Lq, nh, nh, D, H, past, None, 1, pool2.data_ptr(),
           bt.data_ptr(), pps, 1, None, 0, st())
    torch.cuda.synchronize()
    tol = 2 ** -6 * q1.float().abs().max()
    assert (q1.float() - q2.float()).abs().max() <= tol
    assert (pool1.float() - pool2.float()).abs().max() <= tol
    assert (pool2 != 0).any()


@pytest.mark.parametrize('top_p,temp', [(0.9, 1.0), (0.5, 0.7), (1.0, 1.0), (0.05, 1.3)])
def test_top_p_sample_matches_torch_restatement(dev, top_p, temp):
    L = _mods()
    torch.manual_seed(10)
    R, V = 6, 32064
    lg = (torch.randn(R, V, device=dev) * 4).contiguous()
    u = torch.rand(R, device=dev)
    out = torch.zeros(R, dtype=torch.int32, device=dev)
    tau = torch.zeros(R, device=dev)
    L.call('p3_top_p_sample', lg.data_ptr(), R, V, V, top_p, temp, u.data_ptr(), out.data_ptr(), tau.data_ptr(), None, 0, st())
    torch.cuda.synchronize()
    p = torch.softmax(lg.double() / temp, -1)
    sp, si = p.sort(-1, descending=True)
    cum = sp.cumsum(-1)
    keep_sorted = (cum - sp) < top_p                      # smallest prefix whose mass reaches top_p
    tau_ref = torch.where(keep_sorted, sp, torch.ones_like(sp)).min(-1).values
    for r in range(R):
        if top_p < 1.0:
            assert abs(tau[r].item() - tau_ref[r].item()) <= 2e-6 * max(tau_ref[r].item(), 1e-30) + 1e-12
        else:
            assert tau[r].item() == 0.0
            tau_ref[r] = 0.0
        nucleus = p[r] >= tau_ref[r] * (1 - 1e-6)
        assert nucleus[out[r].item()]
        cdf = (p[r] * nucleus).cumsum(-1)
        want = int(torch.searchsorted(cdf, u[r].double() * cdf[-1], right=True).item())
        got = out[r].item()
        if got != want:                                   # fp32 vs fp64 cdf: only a neighbour inside the nucleus is acceptable
            idx = torch.nonzero(nucleus).flatten().tolist()
            assert abs(idx.index(got) - idx.index(min(want, idx[-1]))) <= 1


# ------------------------------------------------------------------ round 2: fused prefill epilogues of the tensor-core GEMM
def _gemm_fused(dev, x, w, out, epi, **kw):
    import ctypes
    import phi3_b200  # noqa
    from phi3_b200 import _lib
    a = _lib.GemmArgs()
    a.X, a.ldx, a.W, a.ldw, a.out, a.ldo = x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0)
    a.M, a.N, a.K, a.epi, a.impl = x.shape[0], w.shape[0], x.shape[1], epi, 0
    for k, v in kw.items():
        setattr(a, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    _lib.call_struct('p3_gemm_fused', a, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()


@pytest.mark.parametrize('M', [40, 781, 2100])
def test_gemm_fused_rmsnorm_scale_and_sumsq(dev, M):
    """RMSNorm folded into the GEMM (gain in W, per-row rsqrt on the accumulators) == rmsnorm kernel + GEMM up to the
    rounding of the folded weights; the residual epilogue's sum-of-squares partials describe the bf16 values written."""
    import phi3_b200  # noqa
    from phi3_b200 import _lib
    H, N = 384, 768
    g = torch.Generator().manual_seed(M)
    x = (torch.randn(M, H, generator=g) * 2).to(torch.bfloat16).to(dev)
    gain = (1 + 0.1 * torch.randn(H, generator=g)).to(torch.bfloat16).to(dev)
    W = (torch.randn(N, H, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    Wf = (W.float() * gain.float()[None]).to(torch.bfloat16).contiguous()
    ss = torch.empty(M, 1, device=dev)
    _lib.call('p3_row_sumsq', x.data_ptr(), x.stride(0), ss.data_ptr(), M, H, torch.cuda.current_stream().cuda_stream)
    assert torch.allclose(ss[:, 0], x.float().pow(2).sum(1), rtol=1e-5)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    plan = _lib.WeightPlan(Wf)
    _gemm_fused(dev, x, Wf, out, _lib.EPI_NONE, ss_in=ss, n_ss_in=1, eps=1e-5, w_plan=plan.addr)
    xf = x.float()
    xn = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * gain.float()).to(torch.bfloat16).float()
    ref = xn @ W.float().T
    assert ((out.float() - ref).abs().max() / ref.abs().max()).item() < 1e-2
    # residual epilogue + ss_out, then a second normed GEMM fed by those partials
    h0 = torch.randn(M, N, generator=g).to(torch.bfloat16).to(dev)
    h = h0.clone()
    ssp = torch.full((M, N // 32), -1.0, device=dev)
    _gemm_fused(dev, x, W, h, _lib.EPI_RESIDUAL, resid=h, ss_out=ssp)
    assert torch.allclose(ssp.sum(1), h.float().pow(2).sum(1), rtol=1e-4)
    assert ((h.float() - (h0.float() + (x.float() @ W.float().T).to(torch.bfloat16).float())).abs().max() / h.float().abs().max()).item() < 1e-2
    W2 = (torch.randn(256, N, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    out2 = torch.zeros(M, 256, dtype=torch.bfloat16, device=dev)
    _gemm_fused(dev, h, W2, out2, _lib.EPI_NONE, ss_in=ssp, n_ss_in=N // 32, eps=1e-5)
    hf = h.float()
    ref2 = (hf * torch.rsqrt(hf.pow(2).mean(-1, keepdim=True) + 1e-5)) @ W2.float().T
    assert ((out2.float() - ref2).abs().max() / ref2.abs().max()).item() < 1e-2


@pytest.mark.parametrize('B,L,nh,past,row_div,wc', [(2, 40, 4, 0, 1, 1), (1, 781, 32, 0, 1, 1), (3, 70, 4, 64, 1, 1), (6, 5, 4, 130, 3, 0)])
def test_gemm_fused_rope_kvwrite_matches_separate_kernels(dev, B, L, nh, past, row_div, wc):
    """P3_EPI_ROPE_KV (permuted W, rope + paged KV write in the epilogue) == p3_gemm + p3_rope_kvwrite bit for bit"""
    import phi3_b200  # noqa
    from phi3_b200 import _lib
    hd, H = 96, 384
    N, M = 3 * nh * hd, B * L
    g = torch.Generator().manual_seed(B * 100 + L)
    x = torch.randn(M, H, generator=g).to(torch.bfloat16).to(dev)
    W = (torch.randn(N, H, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    n_seq = B // row_div
    S = past + L
    pages = (S + 63) // 64
    cos = torch.randn(n_seq, S, hd // 2, generator=g).to(dev)
    sin = torch.randn(n_seq, S, hd // 2, generator=g).to(dev)
    bt = torch.randperm(n_seq * pages, generator=g).to(torch.int32).reshape(n_seq, pages).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    # reference: plain GEMM then the separate rope + KV-write kernel
    qkv_ref = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    pool_ref = torch.zeros(n_seq * pages, 2, nh, 64, hd, dtype=torch.bfloat16, device=dev)
    _lib.call('p3_gemm', x.data_ptr(), x.stride(0), W.data_ptr(), W.stride(0), None, qkv_ref.data_ptr(), qkv_ref.stride(0), None,
              None, M, N, H, _lib.EPI_NONE, 0, st)
    _lib.call('p3_rope_kvwrite', qkv_ref.data_ptr(), cos.data_ptr(), sin.data_ptr(), cos.stride(0), B, L, nh, nh, hd, past, row_div,
              pool_ref.data_ptr(), bt.data_ptr(), bt.stride(0), wc, None, st)
    # fused: permuted weights
    idx = []
    for h in range(2 * nh):
        for j in range(hd // 32):
            idx += [h * hd + 16 * j + i for i in range(16)] + [h * hd + hd // 2 + 16 * j + i for i in range(16)]
    idx += list(range(2 * nh * hd, N))
    Wp = W[torch.tensor(idx, device=dev)].contiguous()
    qkv = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    pool = torch.zeros_like(pool_ref)
    _gemm_fused(dev, x, Wp, qkv, _lib.EPI_ROPE_KV, cosT=cos, sinT=sin, tab_bstride=cos.stride(0), L=L, n_heads=nh, n_kv=nh, hd=hd,
                past=past, row_div=row_div, write_cache=wc, bt_stride=bt.stride(0), pool=pool, block_table=bt)
    assert torch.equal(qkv, qkv_ref)
    assert torch.equal(pool, pool_ref)
    if not wc:
        assert pool.abs().sum() == 0


@pytest.mark.parametrize('L_all,use_pids', [(200, False), (4300, False), (140, True), (131072, False)])
def test_rope_table_on_device(dev, L_all, use_pids):
    """p3_rope_table vs the reference formula in torch fp32 on the CPU (phi:493-504): same fp32 product, cos / sin within 2 ulp
    of the scale (CUDA cosf/sinf vs the host libm), also at 128K positions (arguments up to 1.3e5 rad)."""
    import math
    import phi3_b200  # noqa
    from phi3_b200 import _lib, configs
    cfg = configs.PHI35_MINI
    hd, half = 96, 48
    sf = math.sqrt(1 + math.log(cfg.max_position_embeddings / cfg.original_max_position_embeddings) / math.log(cfg.original_max_position_embeddings))
    fac = cfg.rope_scaling['long_factor'] if L_all > 4096 else cfg.rope_scaling['short_factor']
    inv_freq = 1.0 / (torch.tensor(fac, dtype=torch.float32) * cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    if use_pids:
        Lp = 100
        pids = torch.stack([torch.arange(Lp), torch.cat([torch.ones(30, dtype=torch.long), torch.arange(70)])])
        ext = pids[:, -1][:, None] + 1 + torch.arange(L_all - Lp)[None, :]
        pos = torch.cat([pids, ext], 1).float()
        pid_dev = pids.to(dev, torch.int32).contiguous()
        Bt = 2
    else:
        pos, pid_dev, Lp, Bt = torch.arange(L_all, dtype=torch.float32)[None], None, 0, 1
    fr = pos[:, :, None] * inv_freq[None, None, :]
    ref_c, ref_s = torch.cos(fr) * sf, torch.sin(fr) * sf
    cos = torch.empty(Bt, L_all, half, device=dev)
    sin = torch.empty(Bt, L_all, half, device=dev)
    ifd = inv_freq.to(dev)
    _lib.call('p3_rope_table', None if pid_dev is None else pid_dev.data_ptr(), 0 if pid_dev is None else pid_dev.stride(0), Lp,
              ifd.data_ptr(), cos.data_ptr(), sin.data_ptr(), Bt, L_all, half, float(sf), torch.cuda.current_stream().cuda_stream)
    assert (cos.cpu() - ref_c).abs().max().item() <= 4e-7 * sf * 2 and (sin.cpu() - ref_s).abs().max().item() <= 4e-7 * sf * 2


@pytest.mark.parametrize('M', [5, 80, 128, 320])
@pytest.mark.parametrize('N,K,epi', [(3072, 3072, 'resid'), (9216, 3072, 'none'), (3072, 8192, 'resid'), (16384, 3072, 'swiglu')])
def test_gemm_split_k_small_m(dev, M, N, K, epi):
    """few output tiles and a long K (small M, K = 8192): the tensor-core GEMM splits K over CTAs (fp32 partials, summed in slice
    order by the last CTA of a tile) so that all SMs stream weights; the other shapes must be unaffected by the workspace. Same result as the unsplit kernel up to fp32 summation order; the arrival
    counters are left at zero; repeated launches are bit-identical (deterministic reduction)."""
    import phi3_b200  # noqa
    from phi3_b200 import _lib
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    W = (torch.randn(N, K, generator=g) * 0.03).to(torch.bfloat16).to(dev)
    ws = torch.zeros(32 << 20, dtype=torch.uint8, device=dev)
    kind = {'resid': _lib.EPI_RESIDUAL, 'none': _lib.EPI_NONE, 'swiglu': _lib.EPI_SWIGLU}[epi]
    No = N // 2 if epi == 'swiglu' else N
    h0 = torch.randn(M, No, generator=g).to(torch.bfloat16).to(dev)

    def run(split):
        out = h0.clone() if epi == 'resid' else torch.zeros(M, No, dtype=torch.bfloat16, device=dev)
        kw = dict(resid=out) if epi == 'resid' else {}
        if split:
            kw.update(splitk_ws=ws, splitk_ws_bytes=ws.numel())
        _gemm_fused(dev, x, W, out, kind, **kw)
        return out
    a, b, c = run(False), run(True), run(True)
    assert torch.equal(b, c)
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0
    assert ((a.float() - b.float()).abs().max() / a.float().abs().max()).item() < 1e-2      # <= 2 bf16 ulps of the largest value (two roundings)
    y = x.float() @ W.float().T
    if epi == 'resid':
        ref = h0.float() + y.to(torch.bfloat16).float()
    elif epi == 'swiglu':
        Wi = W.float()                                            # interleaved layout: per 256 rows [128 gate | 128 up]
        yb = y.to(torch.bfloat16).float().reshape(M, N // 256, 2, 128)
        ref = (torch.nn.functional.silu(yb[:, :, 0]).to(torch.bfloat16).float() * yb[:, :, 1]).reshape(M, No)
    else:
        ref = y
    assert ((b.float() - ref).abs().max() / ref.abs().max()).item() < 1e-2
