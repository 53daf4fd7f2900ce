"""-m gpu: the persistent decode-layer kernel (csrc/decode_mega.cu) through the C ABI.
  * p3_mega_pack is bit-exact against the torch restatement of the stream order (mega.pack_reference);
  * every phase kind alone, and the 4-phase chain of a layer, against torch fp32 on the same inputs
    (tolerance: 1e-2 of the output's max — bf16 inputs, fp32 accumulation, one bf16 rounding at the module boundary);
  * the decode loop through the kernel against the per-matrix skinny path and against the CPU oracle."""
import ctypes as C
import math
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mods():
    import phi3_b200  # noqa
    from phi3_b200 import _lib, mega
    return _lib, mega


def _pack(W, kind, nh=0, nkv=0, hd=0):
    _lib, mega = _mods()
    out = torch.empty(W.numel(), dtype=torch.bfloat16, device=W.device)
    _lib.call('p3_mega_pack', W.data_ptr(), out.data_ptr(), kind, W.shape[0], W.shape[1], nh, nkv, hd,
              torch.cuda.current_stream().cuda_stream)
    return out


@pytest.mark.parametrize('kind,N,K,nh,nkv,hd', [
    (0, 384, 384, 0, 0, 0), (1, 2048, 384, 0, 0, 0), (0, 384, 1024, 0, 0, 0), (2, 1152, 384, 4, 4, 96),
    (3, 32064, 384, 0, 0, 0), (0, 3072, 8192, 0, 0, 0), (2, 9216, 3072, 32, 32, 96), (1, 16384, 3072, 0, 0, 0)])
def test_pack_bit_exact(dev, kind, N, K, nh, nkv, hd):
    _lib, mega = _mods()
    W = torch.randn(N, K, generator=torch.Generator().manual_seed(N + K)).to(torch.bfloat16)
    ref = mega.pack_reference(W, kind, nh, nkv, hd)
    got = _pack(W.to(dev), kind, nh, nkv, hd).cpu()
    assert torch.equal(ref.view(torch.int16), got.view(torch.int16))


class _Runner:
    """builds p3_mega_args for ad-hoc phase chains"""

    def __init__(self, dev, M, n_ctas=None):
        _lib, mega = _mods()
        self._lib, self.mega, self.dev, self.M = _lib, mega, dev, M
        self.n_ctas = n_ctas or _lib.lib().p3_decode_mega_ctas()
        self.sync = torch.zeros(2, dtype=torch.int32, device=dev)
        self.keep = []

    def run(self, phases, rope=None):
        """phases: list of dict(kind, W [N,K] bf16 dev, x, out, norm_w, ss_in, ss_out). rope: dict for QKV"""
        mega = self.mega
        a = mega.MegaArgs()
        nh, nkv, hd = (rope['nh'], rope['nkv'], rope['hd']) if rope else (0, 0, 0)
        sched = mega.build_schedule([(p['kind'], p['W'].shape[0], p['W'].shape[1]) for p in phases], self.n_ctas)
        for i, (p, (off, ids, mx)) in enumerate(zip(phases, sched)):
            ph = mega.MegaPhase()
            wp = _pack(p['W'], p['kind'], nh, nkv, hd)
            off, ids = off.to(self.dev), ids.to(self.dev)
            self.keep += [wp, off, ids]
            ph.wp, ph.kind, ph.N, ph.K, ph.max_tiles_per_cta = wp.data_ptr(), p['kind'], p['W'].shape[0], p['W'].shape[1], mx
            ph.x, ph.ldx = p['x'].data_ptr(), p['x'].stride(0)
            ph.norm_w = p['norm_w'].data_ptr() if p.get('norm_w') is not None else None
            ph.ss_in = p['ss_in'].data_ptr() if p.get('ss_in') is not None else None
            ph.n_ss_in = p['ss_in'].shape[0] if p.get('ss_in') is not None else 0
            ph.ss_out = p['ss_out'].data_ptr() if p.get('ss_out') is not None else None
            ph.out, ph.ldo = p['out'].data_ptr(), p['out'].stride(0)
            ph.cta_off, ph.tile_ids = off.data_ptr(), ids.data_ptr()
            a.ph[i] = ph
        a.n_phases, a.M, a.eps, a.n_ctas, a.sync = len(phases), self.M, 1e-5, self.n_ctas, self.sync.data_ptr()
        if rope:
            a.cosT, a.sinT, a.tab_bstride = rope['cos'].data_ptr(), rope['sin'].data_ptr(), rope['tbs']
            a.n_heads, a.n_kv, a.hd, a.past, a.past_dev = nh, nkv, hd, rope['past'], None
            a.pool, a.block_table, a.bt_stride = rope['pool'].data_ptr(), rope['bt'].data_ptr(), rope['bt'].stride(0)
        self._lib.call_struct('p3_decode_mega', a, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert self.sync.tolist() == [0, 0]                     # the barrier words are left clean


def _bf(t):
    return t.to(torch.bfloat16)


def _r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _relmax(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-9)).item()


@pytest.mark.parametrize('M', [1, 3, 8])
@pytest.mark.parametrize('N,K', [(384, 384), (384, 1024), (3072, 8192), (3072, 3072)])
def test_phase_resid(dev, M, N, K):
    g = torch.Generator().manual_seed(M * 1000 + N + K)
    W = _bf(torch.randn(N, K, generator=g) * 0.05).to(dev)
    x = _bf(torch.randn(M, K, generator=g)).to(dev)
    h0 = _bf(torch.randn(M, N, generator=g)).to(dev)
    h = h0.clone()
    ss = torch.full((N // 16, 16), -1.0, device=dev)
    _Runner(dev, M).run([dict(kind=0, W=W, x=x, out=h, ss_out=ss)])
    y = x.float() @ W.float().T
    ref = _r(h0.float() + _r(y))
    assert _relmax(h, ref) < 1e-2
    # sum-of-squares partials describe exactly the bf16 values that were written
    got_ss = ss[:, :M].sum(0)
    assert torch.allclose(got_ss, h.float().pow(2).sum(1), rtol=1e-4)


@pytest.mark.parametrize('M', [1, 5, 8])
@pytest.mark.parametrize('I,K', [(1024, 384), (8192, 3072)])
def test_phase_swiglu_with_norm(dev, M, I, K):
    g = torch.Generator().manual_seed(M + I + K)
    W = _bf(torch.randn(2 * I, K, generator=g) * 0.05).to(dev)
    x = _bf(torch.randn(M, K, generator=g)).to(dev)
    nw = _bf(1 + 0.1 * torch.randn(K, generator=g)).to(dev)
    out = torch.zeros(M, I, dtype=torch.bfloat16, device=dev)
    # ss_in = NULL: the kernel recomputes the row statistics from x
    _Runner(dev, M).run([dict(kind=1, W=W, x=x, out=out, norm_w=nw)])
    xf = x.float()
    xn = _r(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * nw.float())
    y = _r(xn @ W.float().T)
    ref = _r(_r(torch.nn.functional.silu(y[:, :I])) * y[:, I:])
    assert _relmax(out, ref) < 1e-2


@pytest.mark.parametrize('M', [2, 8])
@pytest.mark.parametrize('V,K', [(32064, 384), (32064, 3072)])
def test_phase_logits(dev, M, V, K):
    g = torch.Generator().manual_seed(M + K)
    W = _bf(torch.randn(V, K, generator=g) * 0.05).to(dev)
    x = _bf(torch.randn(M, K, generator=g)).to(dev)
    out = torch.zeros(M, V, dtype=torch.float32, device=dev)
    _Runner(dev, M).run([dict(kind=3, W=W, x=x, out=out)])
    assert _relmax(out, x.float() @ W.float().T) < 2e-3


@pytest.mark.parametrize('M,nh,K', [(3, 4, 384), (8, 32, 3072)])
def test_phase_qkv_rope_kvwrite(dev, M, nh, K):
    hd, past, n_pages = 96, 70, 3
    g = torch.Generator().manual_seed(M + nh)
    N = 3 * nh * hd
    W = _bf(torch.randn(N, K, generator=g) * 0.05).to(dev)
    x = _bf(torch.randn(M, K, generator=g)).to(dev)
    cos = torch.randn(M, past + 4, hd // 2, generator=g).to(dev)
    sin = torch.randn(M, past + 4, hd // 2, generator=g).to(dev)
    pool = torch.zeros(M * n_pages, 2, nh, 64, hd, dtype=torch.bfloat16, device=dev)
    bt = torch.randperm(M * n_pages, generator=g).to(torch.int32).reshape(M, n_pages).to(dev)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    _Runner(dev, M).run([dict(kind=2, W=W, x=x, out=out)],
                        rope=dict(nh=nh, nkv=nh, hd=hd, cos=cos, sin=sin, tbs=cos.stride(0), past=past, pool=pool, bt=bt))
    y = _r(x.float() @ W.float().T).reshape(M, 3, nh, hd)
    c, s = cos[:, past].float()[:, None, :], sin[:, past].float()[:, None, :]

    def rot(v):
        x1, x2 = v[..., :hd // 2], v[..., hd // 2:]
        return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], -1)
    q, k, v = _r(rot(y[:, 0])), _r(rot(y[:, 1])), y[:, 2]
    ref = torch.stack([q, k, v], 1).reshape(M, N)
    assert _relmax(out, ref) < 1e-2
    for b in range(M):
        page = int(bt[b, past // 64])
        assert torch.equal(pool[page, 0, :, past % 64], out[b].reshape(3, nh, hd)[1])
        assert torch.equal(pool[page, 1, :, past % 64], out[b].reshape(3, nh, hd)[2])
    touched = pool.float().abs().sum((1, 2, 3, 4)) > 0
    assert int(touched.sum()) == M                               # exactly one page per row was written


@pytest.mark.parametrize('M,H,I,nh', [(4, 384, 1024, 4), (8, 3072, 8192, 32)])
@pytest.mark.parametrize('n_ctas', [None, 5])
def test_layer_chain(dev, M, H, I, nh, n_ctas):
    """o_proj(+resid) -> norm+gate_up(+SwiGLU) -> down(+resid) -> norm+logits in ONE launch (3 grid barriers)."""
    if n_ctas == 5 and H > 384:
        pytest.skip('small grid only at small sizes (multi-K-block phases keep <= 8 tiles per CTA)')
    g = torch.Generator().manual_seed(H + M)
    rn = lambda *s, sc=1.0: _bf(torch.randn(*s, generator=g) * sc).to(dev)
    Wo, Wgu, Wd, Wl = rn(H, H, sc=0.03), rn(2 * I, H, sc=0.03), rn(H, I, sc=0.02), rn(640, H, sc=0.05)
    att, h0 = rn(M, H), rn(M, H)
    ln2, lnf = _bf(1 + 0.1 * torch.randn(H, generator=g)).to(dev), _bf(1 + 0.1 * torch.randn(H, generator=g)).to(dev)
    h = h0.clone()
    act = torch.zeros(M, I, dtype=torch.bfloat16, device=dev)
    logits = torch.zeros(M, 640, dtype=torch.float32, device=dev)
    ssA, ssB = torch.zeros(H // 16, 16, device=dev), torch.zeros(H // 16, 16, device=dev)
    _Runner(dev, M, n_ctas).run([dict(kind=0, W=Wo, x=att, out=h, ss_out=ssB),
                                 dict(kind=1, W=Wgu, x=h, out=act, norm_w=ln2, ss_in=ssB),
                                 dict(kind=0, W=Wd, x=act, out=h, ss_out=ssA),
                                 dict(kind=3, W=Wl, x=h, out=logits, norm_w=lnf, ss_in=ssA)])

    def rms(x, w):
        return _r(x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float())
    h1 = _r(h0.float() + _r(att.float() @ Wo.float().T))
    gu = _r(rms(h1, ln2) @ Wgu.float().T)
    a = _r(_r(torch.nn.functional.silu(gu[:, :I])) * gu[:, I:])
    h2 = _r(h1 + _r(a @ Wd.float().T))
    ref = rms(h2, lnf) @ Wl.float().T
    assert _relmax(act, a) < 1.5e-2
    assert _relmax(h, h2) < 1.5e-2
    assert _relmax(logits, ref) < 1.5e-2


def test_repeated_launches_are_deterministic(dev):
    M, H = 8, 384
    g = torch.Generator().manual_seed(1)
    W = _bf(torch.randn(H, H, generator=g) * 0.05).to(dev)
    x = _bf(torch.randn(M, H, generator=g)).to(dev)
    h0 = _bf(torch.randn(M, H, generator=g)).to(dev)
    outs = []
    r = _Runner(dev, M)
    for _ in range(3):
        h = h0.clone()
        ss = torch.zeros(H // 16, 16, device=dev)
        r.run([dict(kind=0, W=W, x=x, out=h, ss_out=ss)])
        outs.append((h.clone(), ss.clone()))
    assert all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs[1:])


def _tiny(layers=2, init='gpt2', **kw):
    import os
    os.environ['P3_MEGA'] = '1'                                  # the kernel is opt-in (model.py)
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights
    from phi3_b200.model import Phi3B200
    from oracle.phi3_oracle import Phi3Oracle
    cfg = configs.tiny(layers=layers, **kw)
    w = weights.random_weights(cfg, seed=0, init=init)
    try:
        m = Phi3B200(cfg, w)
    finally:
        os.environ.pop('P3_MEGA', None)
    return cfg, w, m, Phi3Oracle(cfg, w, prec='b200')


@pytest.mark.parametrize('B,quant', [(1, False), (4, False), (8, False), (3, True)])
def test_decode_loop_through_mega(dev, B, quant):
    """graph-replayed decode through the persistent kernel == eager decode through it == per-matrix skinny path (tokens),
    and its logits match the oracle step by step"""
    # peaked checkpoint: token rollouts of two correct bf16 paths (different rounding points: the skinny path applies the RMSNorm
    # scale to the accumulators, the persistent kernel to X) can only be compared where top-1 margins clear the bf16 noise
    cfg, w, m, o = _tiny(use_quantized_cache=quant, init='peaked')
    assert m.mega is not None
    g = torch.Generator().manual_seed(B)
    L = 150 if quant else 40
    ids = torch.randint(3, 32000, (B, L), generator=g)
    ids[:, 0] = 1
    steps = 10
    lg, c = m(ids, max_tokens=steps + 1, logits_rows='last')
    first = lg[:, -1].argmax(-1)
    hist_graph = m.greedy_decode(first, c, steps).cpu()
    del c
    lg, c = m(ids, max_tokens=steps + 1, logits_rows='last')
    hist_eager = m.greedy_decode(first, c, steps, use_graph=False).cpu()
    del c
    mega, m.mega = m.mega, None                                  # same model through the skinny kernels
    lg, c = m(ids, max_tokens=steps + 1, logits_rows='last')
    hist_skinny = m.greedy_decode(first, c, steps, use_graph=False).cpu()
    m.mega = mega
    del c
    assert torch.equal(hist_graph, hist_eager)
    assert torch.equal(hist_graph, hist_skinny)
    # step-wise logits against the oracle (teacher-forced on the oracle's tokens)
    from phi3_b200.model import DecodeSession
    lo, co = o(ids, max_tokens=steps + 1)
    lg, c = m(ids, max_tokens=steps + 1, logits_rows='last')
    tok = lo[:, -1].argmax(-1)
    ses = DecodeSession(m, tok, c, steps, use_graph=False)
    for i in range(steps):
        lo, co = o(tok[:, None], cache=co)
        ses.tok.copy_(tok.to(dev, torch.int32))
        ses.step()
        torch.cuda.synchronize()
        assert _relmax(ses.mega_ses.logits.cpu(), lo[:, -1]) < 2e-2, i
        tok = lo[:, -1].argmax(-1)


def test_left_padded_rows_through_mega(dev):
    cfg, w, m, o = _tiny()
    g = torch.Generator().manual_seed(9)
    lens = [17, 23, 32]
    n = max(lens)
    ids = torch.zeros(3, n, dtype=torch.long)
    pids = torch.ones(3, n, dtype=torch.long)
    mask = torch.zeros(3, n, dtype=torch.long)
    for b, l in enumerate(lens):
        ids[b, n - l:] = torch.randint(3, 32000, (l,), generator=g)
        ids[b, n - l] = 1
        pids[b, n - l:] = torch.arange(l)
        mask[b, n - l:] = 1
    steps = 6
    lo, co = o(ids, pids=pids, mask=mask, max_tokens=steps + 1)
    lg, c = m(ids, pids=pids, mask=mask, max_tokens=steps + 1, logits_rows='last')
    tok = lo[:, -1].argmax(-1)
    from phi3_b200.model import DecodeSession
    ses = DecodeSession(m, tok, c, steps, use_graph=False)
    for i in range(steps):
        lo, co = o(tok[:, None], cache=co)
        ses.tok.copy_(tok.to(dev, torch.int32))
        ses.step()
        torch.cuda.synchronize()
        assert _relmax(ses.mega_ses.logits.cpu(), lo[:, -1]) < 2e-2, i
        tok = lo[:, -1].argmax(-1)
