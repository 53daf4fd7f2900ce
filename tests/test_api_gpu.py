"""-m gpu: the public API (generate / choose / constrain) and the GPU HD transform against the
oracle drivers / oracle processors. Index work is bit-exact; token outputs are compared on
tiny random-init models where near-ties are rare at this vocabulary/sequence size."""
import json
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _setup(vision=False, init='gpt2', **kw):
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights, api
    from phi3_b200.processor import ByteTokenizer
    from oracle.phi3_oracle import Phi3Oracle
    cfg = configs.tiny(vision=vision, **kw)
    clip = configs.tiny_clip(3) if vision else None
    w = weights.random_weights(cfg, seed=3, clip_cfg=clip, init=init)
    model, proc = api.load(blind_model=not vision, cfg=cfg, weights=w, tokenizer=ByteTokenizer(), clip_cfg=clip,
                           num_crops=4, quantize_cache=bool(kw.get('use_quantized_cache', False)))
    return api, model, proc, Phi3Oracle(model.cfg, w, prec='b200', clip_cfg=clip)


@pytest.mark.parametrize('cfgcase', [(336, 336, 4), (500, 350, 4), (300, 420, 4), (640, 480, 16), (1344, 336, 4), (97, 33, 4)])
def test_hd_transform_bit_exact(dev, cfgcase):
    """uint8 padded image bit-exact vs PIL path; pixel_values equal to the oracle's float64 result rounded to fp32."""
    import phi3_b200  # noqa
    from PIL import Image
    from phi3_b200.processor import Phi3VImageProcessor
    from oracle import processors as op
    w, h, nc = cfgcase
    arr = np.random.RandomState(w + h).randint(0, 256, (h, w, 3), dtype=np.uint8)
    ip = Phi3VImageProcessor(num_crops=nc, device=dev)
    pv, shape, ntok, u8 = ip._one(arr)
    ref_u8 = op.hd_transform_u8(Image.fromarray(arr), nc)
    assert torch.equal(u8.cpu(), torch.from_numpy(ref_u8))
    ref = op.image_processor([Image.fromarray(arr)], num_crops=nc, max_crops=1)
    assert shape == ref['image_sizes'][0] and ntok == ref['num_img_tokens'][0]
    ref_pv = torch.from_numpy(ref['pixel_values'][0].astype(np.float32))
    got = pv.cpu()
    assert got.shape == ref_pv.shape
    assert torch.equal(got[1:], ref_pv[1:])                                  # sub crops: LUT values, exact
    assert (got[0] - ref_pv[0]).abs().max() <= 1e-6                          # global crop: float64 sum order only
    out = ip([arr])
    assert out['pixel_values'].shape[1] == max(nc + 1, got.shape[0])


def test_hd_transform_matches_reference_golden_fixture(dev):
    import phi3_b200  # noqa
    from phi3_b200.processor import Phi3VImageProcessor
    gold = json.load(open(os.path.join(HERE, 'golden', 'processor_golden.json')))
    arr = np.load(os.path.join(HERE, 'golden', 'processor_golden.npz'))
    for tag in 'abc':
        meta = gold[f'pixel_{tag}']
        pv, shape, ntok, _ = Phi3VImageProcessor(num_crops=meta['num_crops'], device=dev)._one(arr[f'img_{tag}'])
        assert shape == meta['image_sizes'] and ntok == meta['num_img_tokens'] and pv.shape[0] == meta['n_used']
        sub = pv[:, :, ::7, ::5].cpu().numpy()
        assert np.array_equal(sub[1:], arr[f'pv_{tag}'][1:])
        assert np.abs(sub[0] - arr[f'pv_{tag}'][0]).max() <= 1e-6
        assert np.allclose(pv.double().sum(dim=(1, 2, 3)).cpu().numpy(), arr[f'pvsum_{tag}'], rtol=0, atol=2e-2)


def test_generate_matches_oracle_rollout(dev):
    api, model, proc, ora = _setup()
    from oracle import drivers
    prompts = ['The quick brown fox', 'Hello', 'A somewhat longer prompt about nothing']
    templ, _ = api._apply_chat_template(prompts, None, False)
    inp = proc(templ)
    ref = drivers.generate_ids(ora, inp, 12)
    hist = api._generate(model, proc, templ, max_tokens=12, verbose=False, stream=False, mute=True, return_tokens=True)
    got = hist.cpu().long()
    assert got.shape == ref.shape
    assert torch.equal(got[:, :3], ref[:, :3])
    assert (got == ref).float().mean() >= 0.9
    txt = api.generate(prompts, preload=(model, proc), max_tokens=6, verbose=False, stream=False)
    assert isinstance(txt, list) and len(txt) == 3
    one = api.generate('Hello', preload=(model, proc), max_tokens=5, verbose=False, stream=True)
    assert isinstance(one, str)
    tps = api.generate(prompts, preload=(model, proc), max_tokens=6, verbose=False, return_tps=True)
    assert len(tps) == 2 and tps[1] > 0
    with pytest.raises(ValueError):
        api.generate(prompts, images=[np.zeros((8, 8, 3), np.uint8)], preload=(model, proc), verbose=False)


def test_choose_matches_oracle(dev):
    api, model, proc, ora = _setup()
    from oracle import drivers
    prompts = ['Which letter? A B C D E', 'Another question entirely', 'x']
    templ, _ = api._apply_chat_template(prompts, None, False)
    opts = proc([f' {c}' for c in 'ABCDE'])['input_ids'][:, -1]
    ref = drivers.choose_ids(ora, proc(templ), opts)
    got = api.choose(prompts, preload=(model, proc), verbose=False)
    assert got == ['ABCDE'[i] for i in ref.tolist()]
    assert api.choose('single', preload=(model, proc), verbose=False) in 'ABCDE'


@pytest.mark.parametrize('use_beam', [False, True])
def test_constrain_matches_oracle(dev, use_beam):
    api, model, proc, ora = _setup()
    from oracle import drivers
    prompts = [api._preprocess(p) for p in api._apply_chat_template(['First question here', 'Second, longer question text here'], None, False)[0]]
    text = ' The answer is'
    ids_c = list(proc.tokenizer.encode(text, add_special_tokens=False)[1:])
    inp = proc(prompts)
    synth, score = drivers.constrain_ids(ora, inp, ids_c, 6, use_beam=use_beam, n_beam=3)
    S = inp['input_ids'].shape[1]
    ref_ids = torch.cat([inp['input_ids'], synth], 1).tolist()
    ref_ids = [(r[:r.index(32007, S)] if 32007 in r[S:] else r) for r in ref_ids]
    ref_ids = [[t for t in r if t not in (0, 1)] for r in ref_ids]
    got = api._constrain(model, proc, prompts, [(6, text)], mute=True, verbose=False, use_beam=use_beam, n_beam=3, return_ids=True)
    assert got[0] == ref_ids
    out = api.constrain(['q one', 'q two'], constraints=[(3, ' The'), 'AB'], preload=(model, proc), verbose=False, use_beam=use_beam)
    assert isinstance(out, list) and len(out) == 2 and all(o.rstrip()[-1] in 'AB' for o in out)


def test_quantized_cache_generate_and_beam_raises(dev):
    api, model, proc, ora = _setup(use_quantized_cache=True)
    txt = api.generate(['abc def ghi jkl mno pqr stu vwx yz ' * 3, 'short'], preload=(model, proc), max_tokens=8, verbose=False, stream=False)
    assert len(txt) == 2
    ids = torch.randint(3, 300, (2, 70))
    lg, cache = model(ids, max_tokens=8)
    with pytest.raises(NotImplementedError):
        model(torch.randint(3, 300, (6, 2)), cache=cache, n_beam=3, advance_offset=0)


def test_vision_generate_end_to_end(dev):
    api, model, proc, ora = _setup(vision=True, init='peaked')     # rollouts are compared token by token: margins must clear bf16 noise
    from oracle import processors as op
    from oracle import drivers
    from PIL import Image
    arr = np.random.RandomState(5).randint(0, 256, (350, 500, 3), dtype=np.uint8)
    prompt, imgs = api._apply_chat_template('What is shown?', [arr], False)
    inp = proc(prompt, imgs)
    ref_img = op.image_processor([Image.fromarray(arr)], num_crops=4)
    ref_inp = op.merge(proc.tokenizer, ref_img, prompt)
    assert inp['input_ids'].tolist() == ref_inp['input_ids'].tolist()           # image-token positions bit-exact
    assert inp['positions'].tolist() == ref_inp['positions'].tolist()
    assert inp['image_sizes'].tolist() == ref_inp['image_sizes'].tolist()
    ref_inp = {k: torch.from_numpy(np.asarray(v)) for k, v in ref_inp.items()}
    ref = drivers.generate_ids(ora, ref_inp, 6)
    hist = api._generate(model, proc, prompt, imgs, max_tokens=6, verbose=False, stream=False, mute=True, return_tokens=True)
    assert torch.equal(hist.cpu().long(), ref)


def test_constrain_beam_with_quantized_cache_extension(dev):
    """BASELINE config 5 semantics (quantize_cache + use_beam), defined by extension: identical
    to the bf16-cache algorithm with the prompt KV replaced by its 4-bit g32 image. n_beam 3 and 4."""
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights, api
    from phi3_b200.processor import ByteTokenizer
    from oracle.phi3_oracle import Phi3Oracle
    from oracle import drivers
    cfg = configs.tiny(use_quantized_cache=True)
    w = weights.random_weights(cfg, seed=3)
    model, proc = api.load(blind_model=True, cfg=cfg, weights=w, tokenizer=ByteTokenizer(), quantize_cache=True,
                           allow_beam_with_quantized_cache=True)
    ora = Phi3Oracle(model.cfg, w, prec='b200')
    prompts = [api._preprocess(p) for p in api._apply_chat_template(
        ['A question that is long enough to fill more than one sixty-four token page of the cache, yes indeed it is',
         'Another long question, also spanning more than a single page of the key value cache for sure, really'], None, False)[0]]
    text = ' The answer is'
    ids_c = list(proc.tokenizer.encode(text, add_special_tokens=False)[1:])
    for nb in (3, 4):
        inp = proc(prompts)
        S = inp['input_ids'].shape[1]
        assert S > 64
        synth, _ = drivers.constrain_ids(ora, inp, ids_c, 4, use_beam=True, n_beam=nb)
        ref_ids = torch.cat([inp['input_ids'], synth], 1).tolist()
        ref_ids = [(r[:r.index(32007, S)] if 32007 in r[S:] else r) for r in ref_ids]
        ref_ids = [[t for t in r if t not in (0, 1)] for r in ref_ids]
        got = api._constrain(model, proc, prompts, [(4, text)], mute=True, verbose=False, use_beam=True, n_beam=nb, return_ids=True)
        assert got[0] == ref_ids


def test_generate_top_p_extension(dev):
    api, model, proc, ora = _setup()
    kw = dict(max_tokens=10, verbose=False, stream=False, mute=True, return_tokens=True)
    p = api._apply_chat_template(['hello there', 'general kenobi'], None, False)[0]
    a = api._generate(model, proc, p, top_p=0.9, temperature=1.5, seed=1, **kw).cpu()
    b = api._generate(model, proc, p, top_p=0.9, temperature=1.5, seed=1, **kw).cpu()
    c = api._generate(model, proc, p, top_p=0.9, temperature=1.5, seed=2, **kw).cpu()
    g = api._generate(model, proc, p, **kw).cpu()
    tiny_p = api._generate(model, proc, p, top_p=1e-6, seed=5, **kw).cpu()
    assert torch.equal(a, b) and not torch.equal(a, c)          # seeded, reproducible
    assert torch.equal(tiny_p, g)                               # top_p -> 0 degenerates to greedy


def test_api_edge_cases(dev):
    api, model, proc, ora = _setup()
    # single-token prompt, max_tokens = 1 (no decode loop), B = 1 str in -> str out
    out = api.generate('x', preload=(model, proc), max_tokens=1, verbose=False, stream=False, apply_chat_template=False)
    assert isinstance(out, list) and len(out) == 1      # reference: non-streamed output is always batch_decode's list (pv:72-77)
    out = api.generate('x', preload=(model, proc), max_tokens=2, verbose=False, stream=True, apply_chat_template=False)
    assert isinstance(out, str)
    # constrain with max_new = 0 appends exactly the constraint
    out = api.constrain('Question?', constraints=[(0, ' The')], preload=(model, proc), verbose=False)
    assert out.endswith(' The') or ' The' in out
    # 16 prompts (largest skinny batch) and ragged lengths
    prompts = ['p' * (i + 1) for i in range(16)]
    outs = api.generate(prompts, preload=(model, proc), max_tokens=4, verbose=False, stream=False)
    assert len(outs) == 16
    # 17 prompts (decode falls back to the tensor-core GEMM path, no graph)
    outs = api.generate(prompts + ['q'], preload=(model, proc), max_tokens=3, verbose=False, stream=False)
    assert len(outs) == 17
    ch = api.choose(prompts[:3], choices='AB', preload=(model, proc), verbose=False)
    assert all(c in 'AB' for c in ch)


def test_early_stop_and_streaming_paths(dev, capsys):
    """LogitStopper (pv:79-104) runs un-graphed B=1 steps; Streamer prints incrementally (pv:52-65)."""
    api, model, proc, ora = _setup()
    out = api.generate('Tell me something', preload=(model, proc), max_tokens=12, verbose=True, stream=True, early_stop=3)
    assert isinstance(out, str)
    printed = capsys.readouterr().out
    assert '*** Prompt ***' in printed and 'tokens-per-sec' in printed
    ref = api.generate('Tell me something', preload=(model, proc), max_tokens=12, verbose=False, stream=True)
    assert out == ref[:len(out)] or ref == out[:len(ref)]          # early stop only truncates the greedy continuation


def test_load_from_safetensors_dir(dev, tmp_path):
    """N2 (real-weight loading path): HF-layout safetensors incl. the [O,I,kh,kw] patch-embedding transpose (pv:371-374)."""
    import phi3_b200  # noqa
    from safetensors.torch import save_file
    from phi3_b200 import configs, weights, api
    from phi3_b200.processor import ByteTokenizer
    cfg = configs.tiny(vision=True)
    clip = configs.tiny_clip(3)
    w = weights.random_weights(cfg, seed=5, clip_cfg=clip)
    hf = dict(w)
    key = 'model.vision_embed_tokens.img_processor.vision_model.embeddings.patch_embedding.weight'
    hf[key] = w[key].permute(0, 3, 1, 2).contiguous()               # HF stores [O,I,kh,kw]
    half = len(hf) // 2
    items = list(hf.items())
    save_file({k: v.contiguous() for k, v in items[:half]}, str(tmp_path / 'model-00001-of-00002.safetensors'))
    save_file({k: v.contiguous() for k, v in items[half:]}, str(tmp_path / 'model-00002-of-00002.safetensors'))
    m1, p1 = api.load(cfg=cfg, weights=str(tmp_path), tokenizer=ByteTokenizer(), clip_cfg=clip, num_crops=4)
    m2, p2 = api.load(cfg=cfg, weights=w, tokenizer=ByteTokenizer(), clip_cfg=clip, num_crops=4)
    img = np.random.RandomState(1).randint(0, 256, (300, 400, 3), dtype=np.uint8)
    prompt, imgs = api._apply_chat_template('hi', [img], False)
    a = m1(**p1(prompt, imgs), max_tokens=2)[0]
    b = m2(**p2(prompt, imgs), max_tokens=2)[0]
    assert torch.equal(a, b)


def test_quantize_model_flag_generates_like_the_quantised_oracle(dev):
    """load(quantize_model=True) (pv:1279, 264): weights are quantised at load; generate() rolls out the same tokens as
    the oracle running on its own restatement of nn.quantize(model, 64, 4)."""
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights, api
    from phi3_b200.processor import ByteTokenizer
    from oracle.phi3_oracle import Phi3Oracle, quantize_model_weights
    from oracle import drivers
    cfg = configs.tiny(vision=False)
    w = weights.random_weights(cfg, seed=3)
    model, proc = api.load(blind_model=True, quantize_model=True, cfg=cfg, weights=w, tokenizer=ByteTokenizer())
    assert model.quantize_model
    ora = Phi3Oracle(model.cfg, quantize_model_weights(w), prec='b200')
    prompts = ['The quick brown fox', 'Hello', 'A somewhat longer prompt about nothing']
    templ, _ = api._apply_chat_template(prompts, None, False)
    ref = drivers.generate_ids(ora, proc(templ), 12)
    got = api._generate(model, proc, templ, max_tokens=12, verbose=False, stream=False, mute=True, return_tokens=True).cpu().long()
    assert got.shape == ref.shape and torch.equal(got[:, :2], ref[:, :2])
    assert (got == ref).float().mean() >= 0.85
    txt = api.generate(prompts, preload=(model, proc), max_tokens=6, verbose=False, stream=False)
    assert isinstance(txt, list) and len(txt) == 3


@pytest.mark.parametrize('layers', [1, [0, 2]])
def test_lora_adapter_folded_at_load_matches_loralinear_oracle(dev, layers, tmp_path):
    """use_adapter (pv:266-271, phi:84-133): the product folds scale*alpha/r * (A B)^T into the bf16 weights; the oracle
    evaluates LoRALinear op for op. Also checks the adapter_path route (adapter_config.json + adapters.safetensors)."""
    import json
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights, api
    from phi3_b200.processor import ByteTokenizer
    from oracle.phi3_oracle import Phi3Oracle, lora_modules
    from safetensors.torch import save_file
    cfg = configs.tiny(vision=False, layers=3)
    w = weights.random_weights(cfg, seed=3)
    g = torch.Generator().manual_seed(11)
    targets, r = ['self_attn.qkv_proj', 'mlp.down_proj'], 4
    lcfg = {'model_path': 'x', 'adapter_path': 'y', 'lora_layers': layers, 'lora_targets': targets,
            'lora_parameters': {'rank': r, 'alpha': 8, 'dropout': 0.0, 'scale': 2.0}}
    idx = list(range(3))[-layers:] if isinstance(layers, int) else layers
    lw = {}
    for i in idx:
        for t in targets:
            out_d, in_d = w[f'model.layers.{i}.{t}.weight'].shape
            lw[f'model.layers.{i}.{t}.lora_a'] = (torch.rand(in_d, r, generator=g) * 2 - 1) / in_d ** 0.5
            lw[f'model.layers.{i}.{t}.lora_b'] = torch.randn(r, out_d, generator=g) * 0.1
    json.dump(lcfg, open(tmp_path / 'adapter_config.json', 'w'))
    save_file(lw, str(tmp_path / 'adapters.safetensors'))
    m1, p1 = api.load(blind_model=True, use_adapter=True, adapter={'config': lcfg, 'weights': lw}, cfg=cfg, weights=w, tokenizer=ByteTokenizer())
    m2, _ = api.load(blind_model=True, use_adapter=True, adapter_path=str(tmp_path), cfg=cfg, weights=w, tokenizer=ByteTokenizer())
    m0, _ = api.load(blind_model=True, cfg=cfg, weights=w, tokenizer=ByteTokenizer())
    ora = Phi3Oracle(m1.cfg, w, prec='b200', lora=lora_modules(lcfg, lw, 3))
    ids = torch.randint(3, 32000, (3, 24), generator=g); ids[:, 0] = 1
    lo, co = ora(ids, max_tokens=4)
    l1, c1 = m1(ids, max_tokens=4)
    l2, _ = m2(ids, max_tokens=4)
    l0, _ = m0(ids, max_tokens=4)
    rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max()).item()
    assert rel(l1, lo) < 2e-2
    assert torch.equal(l1, l2)
    assert rel(l0, lo) > 3e-2                               # the adapter changes the function
    tok = lo[:, -1].argmax(-1)
    lo, co = ora(tok[:, None], cache=co)
    l1, c1 = m1(tok[:, None], cache=c1)
    assert rel(l1, lo) < 2e-2


def test_quantize_model_and_quantize_cache_together(dev):
    """both 4-bit flags at once (the reference's smallest-memory configuration): 4-bit g64 weights at decode through the W4
    skinny GEMMs while the prompt KV is read from the 4-bit g32 cache; logits stay within tolerance of the oracle that
    applies the same two quantisers."""
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights, api
    from phi3_b200.processor import ByteTokenizer
    from oracle.phi3_oracle import Phi3Oracle, quantize_model_weights
    cfg = configs.tiny(vision=False, use_quantized_cache=True)
    w = weights.random_weights(cfg, seed=5)
    model, proc = api.load(blind_model=True, quantize_model=True, quantize_cache=True, cfg=cfg, weights=w, tokenizer=ByteTokenizer())
    assert model.quantize_model and model.use_quantized_cache
    ora = Phi3Oracle(model.cfg, quantize_model_weights(w), prec='b200')
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(3, 32000, (2, 70), generator=g); ids[:, 0] = 1
    lo, co = ora(ids, max_tokens=6)
    lg, cg = model(ids, max_tokens=6)
    rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max()).item()
    assert rel(lg, lo) < 2e-2
    tok = lo[:, -1].argmax(-1)
    for _ in range(3):
        lo, co = ora(tok[:, None], cache=co)
        lg, cg = model(tok[:, None], cache=cg)
        assert rel(lg, lo) < 3e-2
        tok = lo[:, -1].argmax(-1)
    txt = api.generate(['abc def ghi', 'short'], preload=(model, proc), max_tokens=6, verbose=False, stream=False)
    assert len(txt) == 2


def test_generate_batch_equals_separate_batch1_calls(dev):
    """generate_batch (extension for BASELINE config 3): N image+text prompts as one left-padded batch give, row by row, the
    tokens of N separate batch-1 generate() calls (the only way the reference can run them, pv:377-378)."""
    api, model, proc, ora = _setup(vision=True, init='peaked')       # peaked logits: token equality is not a coin flip
    rs = np.random.RandomState(11)
    imgs = [rs.randint(0, 256, (350, 500, 3), dtype=np.uint8), rs.randint(0, 256, (400, 400, 3), dtype=np.uint8), None]
    prompts = ['What is shown here?', 'Describe the second picture in a few more words.', 'No image for this one']
    hist = api.generate_batch(prompts, imgs, preload=(model, proc), max_tokens=6, return_tokens=True).cpu()
    assert hist.shape == (3, 6)
    for i, (p, im) in enumerate(zip(prompts, imgs)):
        t, ims = api._apply_chat_template(p, None if im is None else [im], False)
        one = api._generate(model, proc, t, ims, max_tokens=6, verbose=False, stream=False, mute=True, return_tokens=True).cpu()
        assert torch.equal(one[0], hist[i]), (i, one.tolist(), hist[i].tolist())
    txt = api.generate_batch(prompts, imgs, preload=(model, proc), max_tokens=4)
    assert isinstance(txt, list) and len(txt) == 3 and all(isinstance(t, str) for t in txt)
    with pytest.raises(ValueError):
        api.generate_batch(prompts, imgs[:2], preload=(model, proc), max_tokens=4)


def test_dp_generate_with_images_single_process(dev):
    from phi3_b200 import parallel
    api, model, proc, ora = _setup(vision=True, init='peaked')
    rs = np.random.RandomState(12)
    imgs = [rs.randint(0, 256, (336, 336, 3), dtype=np.uint8) for _ in range(3)]
    prompts = ['one', 'two two', 'three three three']
    out = parallel.dp_generate(model, proc, prompts, images=imgs, max_tokens=4, apply_chat_template=True)
    ref = api.generate_batch(prompts, imgs, preload=(model, proc), max_tokens=4)
    assert out == ref and model.force_long_rope is None


def test_short_left_padded_first_call_masks_pad_keys(dev):
    """a left-padded batch with L <= 16 takes the decode-attention kernel on its FIRST call (past == 0); the
    present tile must mask the pad keys like Mask4D does (phi:557-559)."""
    api, model, proc, ora = _setup()
    lens, n = [5, 9, 12], 12
    g = torch.Generator().manual_seed(3)
    ids = torch.zeros(3, n, dtype=torch.long)
    pids = torch.ones(3, n, dtype=torch.long)
    mask = torch.zeros(3, n, dtype=torch.long)
    for b, l in enumerate(lens):
        ids[b, n - l:] = torch.randint(3, 32000, (l,), generator=g)
        ids[b, n - l] = 1
        pids[b, n - l:] = torch.arange(l)
        mask[b, n - l:] = 1
    lo, co = ora(ids, pids=pids, mask=mask, max_tokens=3)
    lg, cg = model(ids, pids=pids, mask=mask, max_tokens=3)
    m = mask.bool()
    rel = ((lg.cpu().float()[m] - lo[m]).abs().max() / lo[m].abs().max()).item()
    assert rel < 2e-2, rel
    tok = lo[:, -1].argmax(-1)
    lo2, _ = ora(tok[:, None], cache=co)
    lg2, _ = model(tok[:, None], cache=cg)
    assert ((lg2.cpu().float() - lo2).abs().max() / lo2.abs().max()).item() < 2e-2


def test_kv_cache_handle_is_a_finite_sequence(dev):
    api, model, proc, ora = _setup()
    lg, c = model(torch.randint(3, 300, (1, 20)), max_tokens=2)
    assert len(list(c)) == model.cfg.num_hidden_layers and c[0].offset == 20
    with pytest.raises(IndexError):
        c[model.cfg.num_hidden_layers]


def test_token_stopper_and_row_truncation_match_oracle(dev):
    """A4: rows that emit EOS (32007) at different steps. The reference stops when every row has emitted EOS once
    (TokenStopper pv:106-117) and cuts each row after its first EOS (Streamer.end pv:73). The checkpoint is the peaked one with
    the next-token permutation patched so that row 0 emits EOS as its 4th token and row 1 as its 8th."""
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights, api
    from phi3_b200.processor import ByteTokenizer
    from oracle.phi3_oracle import Phi3Oracle
    from oracle import drivers
    cfg = configs.tiny()
    w = weights.random_weights(cfg, seed=5, init='peaked')
    emb = w['model.embed_tokens.weight'].float()
    lm = w['lm_head.weight'].clone()
    # (every token has ONE lm_head row, i.e. one predecessor: row 1's chain joins row 0's at the token 'a')
    chains = [[ord('a') + 3, 400, 401, 402, 32007, 403], [ord('b') + 3, 500, 501, 502, ord('a') + 3]]
    for ch in chains:                                              # after token ch[i] the model predicts ch[i+1]
        for a, b in zip(ch[:-1], ch[1:]):
            target = (weights.PEAK_ALPHA * emb[a]).to(lm.dtype)
            lm[(lm == target).all(1)] = 0                           # the permutation's own successor of `a` would tie with b
            lm[b] = target
    w['lm_head.weight'] = lm
    model, proc = api.load(blind_model=True, cfg=cfg, weights=w, tokenizer=ByteTokenizer())
    ora = Phi3Oracle(model.cfg, w, prec='b200')
    prompts = ['xyz a', 'xyz b']                                  # equal lengths: no padding ambiguity
    inp = proc(prompts)
    ref = drivers.generate_ids(ora, inp, 12)
    assert ref.shape[1] == 8 and ref[0, 3] == 32007 and ref[1, 7] == 32007 and (ref[1, :7] != 32007).all()      # the oracle loop stopped when both rows were done
    hist = api._generate(model, proc, prompts, None, max_tokens=12, verbose=False, stream=False, mute=True, return_tokens=True,
                         eos_check_every=1).cpu().long()
    assert hist.shape == ref.shape and torch.equal(hist, ref)
    txt = api._generate(model, proc, prompts, None, max_tokens=12, verbose=False, stream=False, mute=True)
    assert txt == proc.tokenizer.batch_decode(drivers.truncate_rows(ref))
    # polling EOS every 16 steps (the default) decodes past the stop point but returns the same truncated text
    txt16 = api._generate(model, proc, prompts, None, max_tokens=12, verbose=False, stream=False, mute=True, eos_check_every=16)
    assert txt16 == txt


@pytest.mark.parametrize('use_beam', [False, True])
def test_constrain_graph_replay_equals_eager(dev, use_beam):
    """the CUDA-graph constrained step (one replay per forward, cache offset in a device int) == the eager step, ids exact"""
    api, model, proc, ora = _setup(init='peaked')
    prompts = [api._preprocess(p) for p in api._apply_chat_template(['First question here', 'Second, longer question text here', 'q3'], None, False)[0]]
    kw = dict(mute=True, verbose=False, use_beam=use_beam, n_beam=3, return_ids=True)
    a = api._constrain(model, proc, prompts, [(9, ' The'), (5, ' answer')], use_graph=True, alive_check_every=4, **kw)
    b = api._constrain(model, proc, prompts, [(9, ' The'), (5, ' answer')], use_graph=False, **kw)
    assert a == b


def _lora_fixture(cfg, w, seed, layers, targets=('self_attn.qkv_proj', 'mlp.gate_up_proj'), r=4):
    g = torch.Generator().manual_seed(seed)
    n = cfg.num_hidden_layers
    lcfg = {'model_path': 'x', 'adapter_path': 'y', 'lora_layers': layers, 'lora_targets': list(targets),
            'lora_parameters': {'rank': r, 'alpha': 8, 'dropout': 0.0, 'scale': 2.0}}
    idx = list(range(n))[-layers:] if isinstance(layers, int) else layers
    lw = {}
    for i in idx:
        for t in targets:
            out_d, in_d = w[f'model.layers.{i}.{t}.weight'].shape
            lw[f'model.layers.{i}.{t}.lora_a'] = (torch.rand(in_d, r, generator=g) * 2 - 1) / in_d ** 0.5
            lw[f'model.layers.{i}.{t}.lora_b'] = torch.randn(r, out_d, generator=g) * 0.1
    return {'config': lcfg, 'weights': lw}


def test_set_adapter_swaps_in_place_without_reload(dev):
    """N4: adapters are swapped / removed on a loaded model (the reference reloads, pv:266-271). After every swap the logits —
    prefill (fused copies) and decode (skinny stream, captured graph of a recycled slab) — equal a fresh load with that adapter."""
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights, api
    from phi3_b200.processor import ByteTokenizer
    cfg = configs.tiny(vision=False, layers=3)
    w = weights.random_weights(cfg, seed=3)
    ad_a, ad_b = _lora_fixture(cfg, w, 11, 2), _lora_fixture(cfg, w, 12, [0, 2], targets=('self_attn.qkv_proj', 'mlp.down_proj'))
    kw = dict(blind_model=True, cfg=cfg, weights=w, tokenizer=ByteTokenizer())
    live, _ = api.load(use_adapter=True, adapter=ad_a, **kw)
    ids = torch.randint(3, 32000, (3, 40), generator=torch.Generator().manual_seed(1))
    ids[:, 0] = 1

    def run(m):
        lg, c = m(ids, max_tokens=6, logits_rows='last')
        hist = m.greedy_decode(lg[:, -1].argmax(-1), c, 5)
        c.release()
        return lg.clone(), hist

    for ad in (ad_a, ad_b, None, ad_a):
        if ad is not ad_a or ad is None:
            pass
        api.set_adapter(live, ad)
        fresh, _ = api.load(use_adapter=ad is not None, adapter=ad, **kw)
        l1, h1 = run(live)
        l2, h2 = run(fresh)
        assert torch.equal(l1, l2) and torch.equal(h1, h2)
    base, _ = api.load(**kw)
    assert not torch.equal(run(base)[0], run(live)[0])


def test_lora_over_quantized_model(dev):
    """use_adapter + quantize_model: LoRALinear over QuantizedLinear (phi:94-95): base = the 4-bit image of the weight, plus the
    low-rank update; untouched matrices keep streaming as 4-bit codes. Against the oracle: LoRALinear evaluated op for op over
    the quantised weights."""
    import phi3_b200  # noqa
    from phi3_b200 import configs, weights, api
    from phi3_b200.processor import ByteTokenizer
    from oracle.phi3_oracle import Phi3Oracle, lora_modules, quantize_model_weights
    cfg = configs.tiny(vision=False, layers=3)
    w = weights.random_weights(cfg, seed=3)
    ad = _lora_fixture(cfg, w, 21, 2, targets=('self_attn.qkv_proj',))
    m, _ = api.load(blind_model=True, quantize_model=True, use_adapter=True, adapter=ad, cfg=cfg, weights=w, tokenizer=ByteTokenizer())
    assert len(m._w4) > 0 and m.layers[2]['qkv'].data_ptr() not in m._w4 and m.layers[0]['qkv'].data_ptr() in m._w4
    ora = Phi3Oracle(m.cfg, quantize_model_weights(w), prec='b200', lora=lora_modules(ad['config'], ad['weights'], 3))
    ids = torch.randint(3, 32000, (2, 30), generator=torch.Generator().manual_seed(2))
    ids[:, 0] = 1
    lo, co = ora(ids, max_tokens=4)
    lg, cg = m(ids, max_tokens=4)
    rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max()).item()
    assert rel(lg, lo) < 2e-2
    tok = lo[:, -1].argmax(-1)
    lo, co = ora(tok[:, None], cache=co)
    lg, cg = m(tok[:, None], cache=cg)
    assert rel(lg, lo) < 2e-2


def test_server_batches_concurrent_requests_through_the_model(dev):
    """srv:1-38 through the B200 path: concurrent HTTP clients are served by batched generate_batch calls and each gets exactly
    what a direct call with its own prompts returns."""
    import threading
    import urllib.request
    api, model, proc, _ = _setup()
    from phi3_b200 import server
    pre = (model, proc)
    httpd = server.serve(server.model_generate_fn(pre), port=0, max_batch=8, window_ms=50, host='127.0.0.1',
                         group_key=server.rope_side_key(pre))
    port = httpd.server_address[1]
    threading.Thread(target=httpd.serve_forever, daemon=True).start()

    def post(payload):
        req = urllib.request.Request(f'http://127.0.0.1:{port}/v1/completions', data=json.dumps(payload).encode(),
                                     headers={'Content-Type': 'application/json'})
        with urllib.request.urlopen(req, timeout=120) as r:
            return json.loads(r.read().decode())['responses']
    try:
        prompts = ['Hello there', 'What is the capital of France?', 'abc', 'Tell me a story about a dog']
        want = [api.generate_batch([p], preload=pre, max_tokens=6)[0] for p in prompts]
        got = {}
        th = [threading.Thread(target=lambda i=i: got.__setitem__(i, post({'prompt': prompts[i], 'max_tokens': 6}))) for i in range(4)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert [got[i][0] for i in range(4)] == want
        assert httpd.batcher.calls < 4
        assert post({'prompt': prompts[:2], 'max_tokens': 6}) == want[:2]
    finally:
        httpd.shutdown()
        httpd.batcher.close()
