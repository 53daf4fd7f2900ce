"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs
and random-init weights. Tolerance (BASELINE.json north_star): logits within 2e-2 relative
(max abs error / max abs logit), teacher-forced greedy tokens agree on >= 99 % of positions."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _setup(vision=False, layers=2, seed=0, **kw):
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights
    from phi3_b200.model import Phi3B200
    from oracle.phi3_oracle import Phi3Oracle
    cfg = configs.tiny(vision=vision, layers=layers, **kw)
    clip = configs.tiny_clip(3) if vision else None
    w = weights.random_weights(cfg, seed=seed, clip_cfg=clip)
    return cfg, w, Phi3B200(cfg, w, clip_cfg=clip), Phi3Oracle(cfg, w, prec='b200', clip_cfg=clip)


def rel(a, b):
    return ((a.float().cpu() - b.float()).abs().max() / b.float().abs().max()).item()


def _ids(B, L, seed=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 32000, (B, L), generator=g)
    ids[:, 0] = 1
    return ids


def test_prefill_and_decode_equal_length(dev):
    cfg, w, m, o = _setup()
    ids = _ids(4, 32)
    lo, co = o(ids, max_tokens=8)
    lg, cg = m(ids, max_tokens=8)
    assert rel(lg, lo) < TOL
    tok = lo[:, -1].argmax(-1)
    for _ in range(6):
        lo, co = o(tok[:, None], cache=co)
        lg, cg = m(tok[:, None], cache=cg)
        assert rel(lg, lo) < TOL
        tok = lo[:, -1].argmax(-1)


def test_left_padded_batch(dev):
    from phi3_b200.processor import Phi3FProcessor
    cfg, w, m, o = _setup()

    class Tok:
        def __call__(self, texts):
            class E:
                pass
            e = E()
            g = torch.Generator().manual_seed(5)
            e.input_ids = [[1] + torch.randint(3, 32000, (n - 1,), generator=g).tolist() for n in (17, 23, 29, 32)]
            return e
    inp = Phi3FProcessor(Tok())(['a', 'b', 'c', 'd'])
    lo, co = o(**inp, max_tokens=4)
    lg, cg = m(**inp, max_tokens=4)
    valid = inp['mask'].bool()
    assert ((lg.cpu() - lo).abs()[valid].max() / lo[valid].abs().max()).item() < TOL
    tok = lo[:, -1].argmax(-1)
    for _ in range(3):
        lo, co = o(tok[:, None], cache=co)
        lg, cg = m(tok[:, None], cache=cg)
        assert rel(lg, lo) < TOL
        tok = lo[:, -1].argmax(-1)


def test_long_prefill_then_graph_decode(dev):
    cfg, w, m, o = _setup()
    ids = _ids(2, 200, seed=3)
    n_new = 24
    lo, co = o(ids, max_tokens=n_new)
    lg, cg = m(ids, max_tokens=n_new, logits_rows='last')
    assert rel(lg[:, -1], lo[:, -1]) < TOL
    first = lo[:, -1].argmax(-1)
    # oracle greedy rollout
    toks = [first]
    for _ in range(n_new - 1):
        lo, co = o(toks[-1][:, None], cache=co)
        toks.append(lo[:, -1].argmax(-1))
    ref = torch.stack(toks, 1)
    hist = m.greedy_decode(first.to(dev), cg, n_new - 1).cpu().long()
    agree = (hist == ref).float().mean().item()
    # free-running greedy may diverge after a near-tie; the first tokens must match and most of the rest
    assert torch.equal(hist[:, :4], ref[:, :4])
    assert agree >= 0.8, agree
    assert cg.offset == 200 + n_new - 1


def test_teacher_forced_agreement(dev):
    cfg, w, m, o = _setup(layers=4)
    ids = _ids(4, 48, seed=9)
    n = 40
    lo, co = o(ids, max_tokens=n)
    lg, cg = m(ids, max_tokens=n)
    hit = tot = 0
    tok = lo[:, -1].argmax(-1)
    for _ in range(n - 1):
        lo, co = o(tok[:, None], cache=co)
        lg, cg = m(tok[:, None], cache=cg)
        hit += (lg[:, -1].argmax(-1).cpu() == lo[:, -1].argmax(-1)).sum().item()
        tot += tok.numel()
        tok = lo[:, -1].argmax(-1)
    assert hit / tot >= 0.99, (hit, tot)


def test_peek_and_beam_protocol(dev):
    """advance_offset / n_beam semantics of phi.py:523-527, 589-591 used by constrain."""
    cfg, w, m, o = _setup()
    ids = _ids(3, 40, seed=4)
    lo, co = o(ids, max_tokens=20)
    lg, cg = m(ids, max_tokens=20)
    cons = torch.tensor([[11, 12, 13]]).repeat(3, 1)
    lo1, _ = o(cons, cache=co, advance_offset=0)
    lg1, _ = m(cons, cache=cg, advance_offset=0)
    assert rel(lg1, lo1) < TOL and cg.offset == 40 and co[0].offset == 40
    step = torch.cat([lo[:, -1].argmax(-1)[:, None], cons], 1)
    lo2, _ = o(step, cache=co, advance_offset=1)
    lg2, _ = m(step, cache=cg, advance_offset=1)
    assert rel(lg2, lo2) < TOL and cg.offset == 41
    beams = torch.cat([torch.randint(3, 32000, (9, 1), generator=torch.Generator().manual_seed(1)), cons.repeat(3, 1)], 1)
    lo3, _ = o(beams, cache=co, n_beam=3, advance_offset=0)
    lg3, _ = m(beams, cache=cg, n_beam=3, advance_offset=0)
    assert rel(lg3, lo3) < TOL and cg.offset == 41


def test_quantized_cache_decode(dev):
    cfg, w, m, o = _setup(use_quantized_cache=True)
    ids = _ids(2, 150, seed=6)
    lo, co = o(ids, max_tokens=8)
    lg, cg = m(ids, max_tokens=8)
    assert rel(lg, lo) < TOL
    assert cg.n_quant == 128
    tok = lo[:, -1].argmax(-1)
    for _ in range(5):
        lo, co = o(tok[:, None], cache=co)
        lg, cg = m(tok[:, None], cache=cg)
        assert rel(lg, lo) < TOL
        tok = lo[:, -1].argmax(-1)


def test_no_cache_choose_path(dev):
    cfg, w, m, o = _setup()
    ids = _ids(3, 21, seed=8)
    lo, _ = o(ids, max_tokens=0)
    lg, c = m(ids, max_tokens=0)
    assert c is None and rel(lg, lo) < TOL


def test_vision_splice(dev):
    cfg, w, m, o = _setup(vision=True)
    g = torch.Generator().manual_seed(2)
    pv = torch.randn(1, 5, 3, 336, 336, generator=g)
    sizes = torch.tensor([[672, 672]])
    n_img = (2 * 2 + 1) * 144 + 1 + (2 + 1) * 12
    ids = torch.cat([torch.tensor([1, 50, 60]), torch.full((n_img,), -1), torch.tensor([1, 70, 80, 90])])[None]
    pos = torch.nonzero(ids < 0)
    lo, _ = o(ids, pixel_values=pv, image_sizes=sizes, positions=pos, max_tokens=2)
    lg, _ = m(ids, pixel_values=pv, image_sizes=sizes, positions=pos, max_tokens=2)
    assert rel(lg, lo) < TOL


# Full-size parity (BASELINE configs 1-5 shapes, >= 99 % greedy agreement and <= 2e-2 relative logits error against the
# oracle's REFERENCE dtype flow, on the peaked parity checkpoint) lives in tests/test_parity_fullsize_gpu.py.


def test_slab_recycling_and_graph_reuse(dev):
    """Same-shape calls reuse the KV slab + captured decode graph; results must not depend on it,
    and a cache the caller still holds must stay intact."""
    cfg, w, m, o = _setup()
    outs = []
    for seed in (21, 22, 21):
        ids = _ids(2, 70, seed=seed)
        lg, c = m(ids, max_tokens=12, logits_rows='last')
        first = lg[:, -1].argmax(-1)
        hist = m.greedy_decode(first, c, 11)
        outs.append(hist.cpu())
        c.release()                             # returns the slab to the pool (`del c` does the same through __del__)
    assert torch.equal(outs[0], outs[2]) and not torch.equal(outs[0], outs[1])
    assert len(m._slabs[(2, 70, 12, False)]) == 1 and m._slabs[(2, 70, 12, False)][0].session is not None
    # held cache is not clobbered by a second cache of the same shape
    ids_a, ids_b = _ids(2, 70, seed=31), _ids(2, 70, seed=32)
    la, ca = m(ids_a, max_tokens=12)
    lb, cb = m(ids_b, max_tokens=12)
    assert ca.pool.data_ptr() != cb.pool.data_ptr()
    ta = la[:, -1].argmax(-1)
    la2, _ = m(ta[:, None], cache=ca)
    lo, co = o(ids_a, max_tokens=12)
    lo2, _ = o(lo[:, -1].argmax(-1)[:, None], cache=co)
    assert rel(la2, lo2) < TOL


def test_chunked_prefill_matches_oracle(dev):
    """Chunked prefill against the paged pool (the path BASELINE config 4 needs) incl. left padding."""
    cfg, w, m, o = _setup()
    m.prefill_chunk = 128
    ids = _ids(2, 300, seed=41)
    ids[1, :37] = 0
    pids = torch.stack([torch.arange(300), torch.cat([torch.ones(37, dtype=torch.long), torch.arange(263)])])
    mask = torch.ones(2, 300, dtype=torch.long)
    mask[1, :37] = 0
    lo, co = o(ids, pids=pids, mask=mask, max_tokens=6)
    lg, cg = m(ids, pids=pids, mask=mask, max_tokens=6, logits_rows='last')
    assert cg.offset == 300 and rel(lg[:, -1], lo[:, -1]) < TOL
    tok = lo[:, -1].argmax(-1)
    for _ in range(3):
        lo, co = o(tok[:, None], cache=co)
        lg, cg = m(tok[:, None], cache=cg)
        assert rel(lg, lo) < TOL
        tok = lo[:, -1].argmax(-1)


def test_long_rope_factors_past_4096(dev):
    """L_all > 4096 switches SuRoPE to the long factors (phi.py:492, H7); chunked prefill + decode."""
    cfg, w, m, o = _setup()
    m.prefill_chunk = 2048
    ids = _ids(1, 4200, seed=43)
    lo, co = o(ids, max_tokens=4)
    lg, cg = m(ids, max_tokens=4, logits_rows='last')
    assert rel(lg[:, -1], lo[:, -1]) < TOL
    tok = lo[:, -1].argmax(-1)
    lo, co = o(tok[:, None], cache=co)
    lg, cg = m(tok[:, None], cache=cg)
    assert rel(lg, lo) < TOL
    # the switch really matters: forcing short factors must change the result
    m.force_long_rope = False
    lg_short, _ = m(ids, max_tokens=4, logits_rows='last')
    m.force_long_rope = None
    assert rel(lg_short[:, -1], lo[:, -1] * 0 + lg[:, -1].cpu()) > 1e-3


@pytest.mark.parametrize('shape', [(1, 1), (1, 16), (1, 17), (16, 1), (2, 8), (3, 63), (2, 64), (2, 65), (5, 129)])
def test_token_count_and_page_boundaries(dev, shape):
    """skinny (<=16 tokens) vs tensor-core GEMM routing, L = 1, page-boundary prompt lengths, decode across a
    page boundary, B up to 16."""
    cfg, w, m, o = _setup()
    B, L = shape
    ids = _ids(B, L, seed=50 + B * 131 + L)
    lo, co = o(ids, max_tokens=4)
    lg, cg = m(ids, max_tokens=4)
    assert rel(lg, lo) < TOL
    tok = lo[:, -1].argmax(-1)
    for _ in range(3):
        lo, co = o(tok[:, None], cache=co)
        lg, cg = m(tok[:, None], cache=cg)
        assert rel(lg, lo) < TOL
        tok = lo[:, -1].argmax(-1)


def test_heavy_left_padding_skips_whole_pages(dev):
    """kv_start beyond one 64-token page: the decode kernel starts at a later tile; pad rows stay finite."""
    cfg, w, m, o = _setup()
    L = 200
    ids = _ids(2, L, seed=61)
    pad = 150
    ids[1, :pad] = 0
    pids = torch.stack([torch.arange(L), torch.cat([torch.ones(pad, dtype=torch.long), torch.arange(L - pad)])])
    mask = torch.ones(2, L, dtype=torch.long)
    mask[1, :pad] = 0
    lo, co = o(ids, pids=pids, mask=mask, max_tokens=5)
    lg, cg = m(ids, pids=pids, mask=mask, max_tokens=5)
    assert torch.isfinite(lg).all()
    valid = mask.bool()
    assert ((lg.cpu() - lo).abs()[valid].max() / lo[valid].abs().max()).item() < TOL
    tok = lo[:, -1].argmax(-1)
    for _ in range(4):
        lo, co = o(tok[:, None], cache=co)
        lg, cg = m(tok[:, None], cache=cg)
        assert rel(lg, lo) < TOL
        tok = lo[:, -1].argmax(-1)


def test_two_images_one_prompt(dev):
    """two images of different crop grids spliced into one row (positions / idx advance, phi.py:412-415)"""
    cfg, w, m, o = _setup(vision=True)
    g = torch.Generator().manual_seed(7)
    pv = torch.zeros(2, 5, 3, 336, 336)
    pv[0] = torch.randn(5, 3, 336, 336, generator=g)                  # 2x2 grid
    pv[1, :3] = torch.randn(3, 3, 336, 336, generator=g)              # 1x2 grid (+2 zero-pad crops)
    sizes = torch.tensor([[672, 672], [336, 672]])
    n1 = (4 + 1) * 144 + 1 + 3 * 12
    n2 = (2 + 1) * 144 + 1 + 2 * 12
    ids = torch.cat([torch.tensor([1, 9]), torch.full((n1,), -1), torch.tensor([1, 7]), torch.full((n2,), -2), torch.tensor([1, 5, 6])])[None]
    pos = torch.nonzero(ids < 0)
    lo, _ = o(ids, pixel_values=pv, image_sizes=sizes, positions=pos, max_tokens=2)
    lg, _ = m(ids, pixel_values=pv, image_sizes=sizes, positions=pos, max_tokens=2)
    assert rel(lg, lo) < TOL


def test_gqa_is_rejected_like_the_reference_graph(dev):
    import phi3_b200  # noqa
    from phi3_b200 import configs
    from phi3_b200.model import Phi3B200
    cfg = configs.tiny(num_key_value_heads=2)
    with pytest.raises(NotImplementedError):
        Phi3B200(cfg, {})


def _setup_q(vision=False, layers=2, seed=0):
    """quantize_model=True on both sides: CUDA model quantises at load (quant.py), the oracle runs the fp32 model over
    its own restatement of nn.quantize(model, 64, 4) (oracle.quantize_model_weights)."""
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights
    from phi3_b200.model import Phi3B200
    from oracle.phi3_oracle import Phi3Oracle, quantize_model_weights
    cfg = configs.tiny(vision=vision, layers=layers)
    clip = configs.tiny_clip(3) if vision else None
    w = weights.random_weights(cfg, seed=seed, clip_cfg=clip)
    return cfg, w, Phi3B200(cfg, w, clip_cfg=clip, quantize_model=True), \
        Phi3Oracle(cfg, quantize_model_weights(w), prec='b200', clip_cfg=clip)


def test_quantize_model_prefill_and_decode(dev):
    """pv:264,291-305: 4-bit g64 weights. Prefill runs the tensor-core GEMMs on the dequantised image, decode streams
    the 4-bit codes (p3_gemm_skinny_w4 / p3_gemm_skinny_qkv_rope_w4); both against the oracle's quantised model."""
    cfg, w, m, o = _setup_q()
    assert m.quantize_model and len(m._w4) == 4 * cfg.num_hidden_layers + 1
    ids = _ids(4, 40)
    lo, co = o(ids, max_tokens=8)
    lg, cg = m(ids, max_tokens=8)
    assert rel(lg, lo) < TOL
    tok = lo[:, -1].argmax(-1)
    for _ in range(6):
        lo, co = o(tok[:, None], cache=co)
        lg, cg = m(tok[:, None], cache=cg)
        assert rel(lg, lo) < TOL
        tok = lo[:, -1].argmax(-1)
    # the quantised model is a different function from the bf16 one (guards against a silently ignored flag)
    _, _, m16, _ = _setup()
    l16, _ = m16(ids, max_tokens=8)
    lq, _ = m(ids, max_tokens=8)
    assert rel(lq, l16.float().cpu()) > 5e-2


def test_quantize_model_graph_decode_matches_stepwise(dev):
    cfg, w, m, o = _setup_q()
    ids = _ids(3, 24, seed=4)
    lg, cg = m(ids, max_tokens=12)
    tok = lg[:, -1].argmax(-1)
    hist = m.greedy_decode(tok, cg, 10, use_graph=True)
    lo, co = o(ids, max_tokens=12)
    t = lo[:, -1].argmax(-1)
    agree, n = 0, 0
    for i in range(10):
        lo, co = o(hist[:, i].cpu().long()[:, None], cache=co)              # teacher-forced on the CUDA tokens
        agree += (lo[:, -1].argmax(-1) == hist[:, i + 1].cpu().long()).sum().item()
        n += hist.shape[0]
    assert agree / n >= 0.9


def test_quantize_model_vision(dev):
    """every CLIP / projector Linear and the position embedding are quantised too (nn.quantize walks the whole model)"""
    cfg, w, m, o = _setup_q(vision=True)
    g = torch.Generator().manual_seed(2)
    pv = torch.randn(1, 5, 3, 336, 336, generator=g)
    sizes = torch.tensor([[672, 672]])
    n_img = (2 * 2 + 1) * 144 + 1 + (2 + 1) * 12
    ids = torch.cat([torch.tensor([1, 50, 60]), torch.full((n_img,), -1), torch.tensor([1, 70, 80, 90])])[None]
    pos = torch.nonzero(ids < 0)
    lo, _ = o(ids, pixel_values=pv, image_sizes=sizes, positions=pos, max_tokens=2)
    lg, _ = m(ids, pixel_values=pv, image_sizes=sizes, positions=pos, max_tokens=2)
    assert rel(lg, lo) < 3e-2
