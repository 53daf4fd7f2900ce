"""-m "not gpu": pins the CPU oracle — against HF transformers (decoder, CLIP tower), against the
reference's own processor code (golden fixtures), and checks internal consistency of the quirk
handling (left pad, quantised cache, beam/peek protocol)."""
import json
import os
import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _tiny(vision=False, **kw):
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights
    cfg = configs.tiny(vision=vision, **kw)
    clip = configs.tiny_clip(3) if vision else None
    return cfg, clip, weights.random_weights(cfg, seed=0, clip_cfg=clip)


def test_decoder_matches_hf_phi3():
    from transformers import Phi3Config, Phi3ForCausalLM
    from oracle.phi3_oracle import Phi3Oracle
    cfg, _, w = _tiny()
    hc = Phi3Config(hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=4,
                    num_key_value_heads=4, intermediate_size=cfg.intermediate_size, vocab_size=cfg.vocab_size,
                    rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=131072,
                    original_max_position_embeddings=4096, pad_token_id=0, attn_implementation='eager',
                    rope_scaling={"type": "longrope", "short_factor": cfg.rope_scaling['short_factor'],
                                  "long_factor": cfg.rope_scaling['long_factor']})
    hf = Phi3ForCausalLM(hc).float().eval()
    missing = hf.load_state_dict({k: v.float() for k, v in w.items()}, strict=False)
    assert not missing.unexpected_keys and all('rotary' in k or 'inv_freq' in k for k in missing.missing_keys)
    ids = torch.randint(3, 32000, (1, 19), generator=torch.Generator().manual_seed(0))
    o = Phi3Oracle(cfg, w, prec='fp32')
    lo, cache = o(ids, max_tokens=3)
    with torch.no_grad():
        lh = hf(ids).logits
    assert (lo - lh).abs().max() < 1e-4
    # incremental decode == full forward (KV cache bookkeeping)
    nxt = lo[:, -1].argmax(-1)[:, None]
    l2, _ = o(nxt, cache=cache)
    with torch.no_grad():
        lh2 = hf(torch.cat([ids, nxt], 1)).logits[:, -1]
    assert (l2[:, -1] - lh2).abs().max() < 1e-4


def test_clip_tower_matches_hf():
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from oracle.phi3_oracle import Phi3Oracle
    cfg, clip, w = _tiny(vision=True)
    hc = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=clip.num_hidden_layers,
                          num_attention_heads=16, image_size=336, patch_size=14, hidden_act='quick_gelu',
                          layer_norm_eps=1e-5, attn_implementation='eager')
    hf = CLIPVisionModel(hc).float().eval()
    P = 'model.vision_embed_tokens.img_processor.'
    sd = {}
    for k, v in w.items():
        if k.startswith(P):
            v = v.float()
            if k.endswith('patch_embedding.weight'):
                v = v.permute(0, 3, 1, 2).contiguous()                 # reference layout [O,kh,kw,I] -> HF [O,I,kh,kw]
            sd[k[len(P):]] = v
    res = hf.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    px = torch.randn(2, 3, 336, 336, generator=torch.Generator().manual_seed(1))
    o = Phi3Oracle(cfg, w, prec='fp32', clip_cfg=clip)
    mine = o.clip(px)
    with torch.no_grad():
        ref = hf(px, output_hidden_states=True).hidden_states[-2][:, 1:]
    assert (mine - ref).abs().max() < 2e-4


def test_left_pad_rows_equal_unpadded():
    """H1/H8: a left-padded row must give the same logits on its valid positions as the same
    prompt run alone... except positions, which the reference shifts: pids restart at 0 after the pad."""
    from oracle.phi3_oracle import Phi3Oracle
    cfg, _, w = _tiny()
    o = Phi3Oracle(cfg, w, prec='fp32')
    g = torch.Generator().manual_seed(3)
    a = torch.randint(3, 32000, (1, 9), generator=g)
    b = torch.randint(3, 32000, (1, 14), generator=g)
    la, _ = o(a, max_tokens=0)
    ids = torch.cat([torch.cat([torch.zeros(1, 5, dtype=torch.long), a], 1), b], 0)
    pids = torch.tensor([[1] * 5 + list(range(9)), list(range(14))])
    mask = torch.tensor([[0] * 5 + [1] * 9, [1] * 14])
    lb, _ = o(ids, pids=pids, mask=mask, max_tokens=0)
    assert (lb[0, 5:] - la[0]).abs().max() < 1e-4
    assert torch.isfinite(lb).all()


def test_quantizer_properties():
    from oracle.phi3_oracle import quantize_q4g32, dequantize_q4g32
    x = torch.randn(6, 96 * 5, generator=torch.Generator().manual_seed(0)) * 3
    for prec in ('ref', 'b200'):
        q, s, b = quantize_q4g32(x, prec)
        d = dequantize_q4g32(q, s, b, x.shape)
        assert q.max() <= 15 and q.min() >= 0
        assert (d - x).abs().max() <= 1.1 * s.abs().reshape(-1).max() + 1e-6   # edge-exact rule can clip the far edge by < 1 step
    # the larger-magnitude edge of every group is reproduced exactly (mlx rule)
    q, s, b = quantize_q4g32(x, 'ref')
    d = dequantize_q4g32(q, s, b, x.shape).reshape(-1, 32)
    xg = x.reshape(-1, 32)
    edge = torch.where(xg.min(1).values.abs() > xg.max(1).values.abs(), xg.min(1).values, xg.max(1).values)
    got = torch.where(xg.min(1).values.abs() > xg.max(1).values.abs(), d.min(1).values, d.max(1).values)
    assert (edge - got).abs().max() < 1e-5


def test_quantized_cache_semantics():
    """phi.py:528-540: prefill sees exact K,V; later steps read the quantised prompt image."""
    from oracle.phi3_oracle import Phi3Oracle
    cfg, _, w = _tiny()
    cq, _, _ = _tiny(use_quantized_cache=True)
    ids = torch.randint(3, 32000, (2, 40), generator=torch.Generator().manual_seed(2))
    a, b = Phi3Oracle(cfg, w, 'fp32'), Phi3Oracle(cq, w, 'fp32')
    la, ca = a(ids, max_tokens=4)
    lb, cb = b(ids, max_tokens=4)
    assert torch.equal(la, lb)
    t = la[:, -1].argmax(-1)[:, None]
    la2, _ = a(t, cache=ca)
    lb2, _ = b(t, cache=cb)
    d = (la2 - lb2).abs().max()
    assert 0 < d < 0.5 * la2.abs().max()


def test_drivers_consistency():
    from oracle.phi3_oracle import Phi3Oracle
    from oracle import drivers
    cfg, _, w = _tiny()
    o = Phi3Oracle(cfg, w, 'fp32')
    ids = torch.randint(3, 32000, (2, 12), generator=torch.Generator().manual_seed(4))
    toks = drivers.generate_ids(o, {'input_ids': ids}, 5)
    assert toks.shape == (2, 5)
    # teacher-forced full forward reproduces the greedy tokens
    full, _ = o(torch.cat([ids, toks[:, :-1]], 1), max_tokens=0)
    assert torch.equal(full[:, 11:].argmax(-1), toks)
    # constrain with max_new=0 returns the constraint itself + EOS; scores are finite
    synth, score = drivers.constrain_ids(o, {'input_ids': ids}, [5, 6, 7], 0)
    assert synth.tolist() == [[5, 6, 7, 32007]] * 2 and torch.isfinite(score).all()
    s_nb, sc_nb = drivers.constrain_ids(o, {'input_ids': ids}, [5, 6, 7], 3, use_beam=False)
    s_b, sc_b = drivers.constrain_ids(o, {'input_ids': ids}, [5, 6, 7], 3, use_beam=True)
    assert s_nb.shape == s_b.shape == (2, 7) and (sc_b >= sc_nb - 1e-6).all()


# ----------------------------------------------------------------------------- processors vs reference goldens
def _gold():
    return (json.load(open(os.path.join(HERE, 'golden', 'processor_golden.json'))),
            np.load(os.path.join(HERE, 'golden', 'processor_golden.npz')))


def test_geometry_matches_reference_goldens():
    import phi3_b200  # noqa
    from phi3_b200.processor import hd_geometry
    gold, _ = _gold()
    n = 0
    for row in gold['geometry']:
        if 'error' in row:
            continue
        g = hd_geometry(row['w'], row['h'], row['num_crops'])
        assert [g['H'], g['W']] == row['image_sizes'], row
        assert g['num_img_tokens'] == row['num_img_tokens'], row
        crops = (g['H'] // 336) * (g['W'] // 336) + 1
        assert max(17, crops) == row['pv_shape'][1]
        n += 1
    assert n >= 20
    # table in SURVEY.md A.1
    assert hd_geometry(1600, 1000, 16)['num_img_tokens'] == 3085
    assert hd_geometry(672, 672, 4)['num_img_tokens'] == 757


def test_tokenize_and_merge_match_reference_goldens():
    import phi3_b200  # noqa
    from phi3_b200.processor import Phi3FProcessor, Phi3VProcessor
    from oracle import processors as op
    from tests.golden.make_golden import FakeTok
    gold, _ = _gold()
    mine = Phi3FProcessor(FakeTok())._tokenize(['abc', 'a', 'hello'])
    ora = op.tokenize(FakeTok(), ['abc', 'a', 'hello'])
    for k in ('input_ids', 'pids', 'mask'):
        assert mine[k].tolist() == gold['tokenize'][k] == ora[k].tolist()
    vp = Phi3VProcessor.__new__(Phi3VProcessor)
    vp.tokenizer = FakeTok()
    fake = {'pixel_values': torch.zeros(2, 1), 'image_sizes': [[336, 336], [336, 336]], 'num_img_tokens': [5, 3]}
    m = vp._merge(fake, gold['merge']['text'])
    o = op.merge(FakeTok(), fake, gold['merge']['text'])
    assert m['input_ids'].tolist() == gold['merge']['input_ids'] == o['input_ids'].tolist()
    assert m['positions'].tolist() == gold['merge']['positions'] == o['positions'].tolist()


def test_oracle_image_processor_matches_reference_goldens():
    from PIL import Image
    from oracle import processors as op
    gold, arr = _gold()
    for tag in 'abc':
        meta = gold[f'pixel_{tag}']
        out = op.image_processor([Image.fromarray(arr[f'img_{tag}'])], num_crops=meta['num_crops'], max_crops=17)
        assert out['image_sizes'][0] == meta['image_sizes'] and out['num_img_tokens'][0] == meta['num_img_tokens']
        pv = out['pixel_values'][0]
        n = meta['n_used']
        assert np.array_equal(pv[:n, :, ::7, ::5].astype(np.float32), arr[f'pv_{tag}'])
        assert np.allclose(pv[:n].sum(axis=(1, 2, 3)), arr[f'pvsum_{tag}'], rtol=0, atol=1e-6)
        assert np.abs(pv[n:]).sum() == arr[f'pvpad_{tag}'][0] == 0


def test_interp336_tables_match_reference():
    import phi3_b200  # noqa
    from phi3_b200.processor import interp336_tables
    from oracle.processors import interp336_weights
    _, arr = _gold()
    for n, live in ((336, 336), (672, 168), (1008, 112), (1344, 84), (1680, 68)):
        w_ref, i_ref = arr[f'i336_w_{n}'], arr[f'i336_i_{n}']
        w_o, i_o = interp336_weights(n)
        assert np.array_equal(w_o, w_ref) and np.array_equal(i_o, i_ref)
        idx, wgt = interp336_tables(n)
        assert np.array_equal(idx, i_ref[:, :2]) and np.array_equal(wgt, w_ref[:, :2])
        assert int((np.abs(w_ref).sum(1) > 0).sum()) == live


def test_pil_coefficient_tables_reproduce_pil_resize():
    """The host-built tables + the kernels' integer arithmetic (restated in numpy) == PIL.Image.resize."""
    import phi3_b200  # noqa
    from PIL import Image
    from phi3_b200.processor import pil_bilinear_coeffs
    rng = np.random.RandomState(1)
    for (w, h, nw, nh) in [(500, 350, 672, 470), (640, 480, 1344, 1008), (1920, 1080, 1680, 945), (200, 300, 336, 504),
                           (336, 336, 336, 336), (97, 33, 672, 228)]:
        img = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.array(Image.fromarray(img).resize([nw, nh], Image.BILINEAR))
        bh, kh, _ = pil_bilinear_coeffs(w, nw)
        bv, kv, _ = pil_bilinear_coeffs(h, nh)

        def pass1d(a, bounds, kk):                                        # along axis 1
            out = np.zeros((a.shape[0], bounds.shape[0], 3), dtype=np.uint8)
            for xx in range(bounds.shape[0]):
                x0, n = bounds[xx]
                acc = (a[:, x0:x0 + n].astype(np.int64) * kk[xx, :n][None, :, None]).sum(1) + (1 << 21)
                out[:, xx] = np.clip(acc >> 22, 0, 255)
            return out
        tmp = pass1d(img, bh, kh)
        got = pass1d(tmp.transpose(1, 0, 2), bv, kv).transpose(1, 0, 2)
        assert np.array_equal(got, ref), (w, h, nw, nh)
