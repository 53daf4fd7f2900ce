"""-m "not gpu": host-side logic — the C-ABI library loads and exports every symbol the header
declares (no compute without a GPU), ctypes signatures match the header, the product never
imports the oracle, DP sharding over world_size 2 (gloo), chat template / stoppers."""
import os
import re
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'phi-3-vision-mlx_b200')


def _header_decls():
    h = open(os.path.join(ROOT, 'include', 'phi3_b200.h')).read()
    h = re.sub(r'/\*.*?\*/', '', h, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(?:int|int64_t|const char\*)\s+(p3_\w+)\s*\(([^;]*?)\)\s*;', h, flags=re.S):
        args = [a for a in m.group(2).split(',') if a.strip() and a.strip() != 'void']
        out[m.group(1)] = len(args)
    return out


def test_library_exports_every_declared_symbol():
    import ctypes
    import phi3_b200  # noqa
    from phi3_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), 'build the extension first (__graft_entry__.build())'
    L = ctypes.CDLL(_lib.LIB_PATH)
    decls = _header_decls()
    assert len(decls) >= 20
    for name in decls:
        assert hasattr(L, name), f'{name} declared in include/phi3_b200.h but not exported'
    assert L.p3_version() == 1


def test_ctypes_signatures_match_header():
    import phi3_b200  # noqa
    from phi3_b200 import _lib
    decls = _header_decls()
    for name, sig in _lib._SIGS.items():
        assert name in decls, name
        assert len(sig) == decls[name], (name, len(sig), decls[name])
    with pytest.raises(TypeError):
        _lib.call('p3_rmsnorm', 0, 0)


def test_bad_arguments_fail_loudly_without_gpu():
    """argument validation happens before any launch, so it is testable on a CPU box"""
    import phi3_b200  # noqa
    from phi3_b200 import _lib
    with pytest.raises(RuntimeError, match='M must be in'):
        _lib.call('p3_gemm_skinny', None, 8, None, 1e-5, None, None, 8, None, 17, 64, 64, 0, None, 0, None, None, 0, None)
    with pytest.raises(RuntimeError, match='head_dim must be 96 or 64'):
        _lib.call('p3_attention_prefill', None, None, None, 8, 8, 8, None, 8, 1, 1, 1, 1, 80, 1.0, 1, 0, None, None, None, 0, 1, None)
    with pytest.raises(RuntimeError, match='n_top'):
        _lib.call('p3_row_stats', None, 1, 8, 8, None, None, None, 9, None, None, 0, None, None, None)


def test_model_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import phi3_b200  # noqa
    from phi3_b200 import configs
    from phi3_b200.model import Phi3B200
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Phi3B200(configs.tiny(), {})


def test_product_never_imports_oracle():
    for fn in os.listdir(PKG):
        if fn.endswith('.py'):
            src = open(os.path.join(PKG, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), fn
    for fn in os.listdir(os.path.join(PKG, 'csrc')):
        if fn.endswith(('.cu', '.cuh')):
            src = open(os.path.join(PKG, 'csrc', fn)).read()
            # comments may cite the oracle's rule; code must not include, link or load anything from oracle/
            assert not re.search(r'#\s*include[^\n]*oracle', src), fn
            assert 'oracle/' not in src and 'liboracle' not in src, fn


def test_chat_template_and_preprocess():
    import phi3_b200  # noqa
    from phi3_b200.api import _apply_chat_template, _preprocess, LogitStopper
    p, im = _apply_chat_template('  hi there ', None, False)
    assert p == '<|user|>\nhi there<|end|>\n<|assistant|>\n' and im is None
    p, _ = _apply_chat_template(['a', 'b'], None, False)
    assert p == ['<|user|>\na<|end|>\n<|assistant|>\n', '<|user|>\nb<|end|>\n<|assistant|>\n']
    p, im = _apply_chat_template('x', [object(), object()], False)
    assert p.startswith('<|user|>\n<|image_1|>\n<|image_2|>\nx<|end|>')
    assert _apply_chat_template('raw', None, False, apply_chat_template=False)[0] == 'raw'
    assert _preprocess('<|user|> a<|end|><|assistant|>') == '<|user|>\na<|end|>\n<|assistant|>'
    ls = LogitStopper(10, 3)
    assert ls.early_stop == 3 and LogitStopper(10, False).early_stop is False and LogitStopper(10, 20).early_stop is False


def test_byte_tokenizer_roundtrip():
    import phi3_b200  # noqa
    from phi3_b200.processor import ByteTokenizer, Phi3FProcessor
    t = ByteTokenizer()
    s = '<|user|>\nWhat is 2+2?<|end|>\n<|assistant|>\n'
    ids = t(s).input_ids
    assert ids[0] == 1 and 32007 in ids and t.decode(ids[1:]) == s
    out = Phi3FProcessor(t)(['ab', 'abcd'])
    assert out['input_ids'].shape == (2, 5) and out['input_ids'][0, :2].tolist() == [0, 0]
    assert out['pids'][0].tolist() == [1, 1, 0, 1, 2] and out['mask'][0].tolist() == [0, 0, 1, 1, 1]


def test_shard_indices_and_rope_switch():
    import phi3_b200  # noqa
    from phi3_b200.parallel import shard_indices, global_rope_switch, dp_map
    sh = shard_indices([5, 9, 1, 7, 3], 2)
    assert sorted(sh[0] + sh[1]) == [0, 1, 2, 3, 4] and sh[0] == [1, 0, 2] and sh[1] == [3, 4]
    assert shard_indices([1, 2], 4) == [[1], [0], [], []]
    assert global_rope_switch([100, 4000], 128) and not global_rope_switch([100, 3000], 128)
    assert dp_map(lambda items, idx: [x * 2 for x in items], [1, 2, 3]) == [2, 4, 6]


_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
import phi3_b200
from phi3_b200.parallel import dp_map
dist.init_process_group('gloo', rank=int(os.environ['RANK']), world_size=2)
items = [f'p{i}' for i in range(7)]
lens = [3, 9, 4, 8, 1, 7, 2]
seen = []
def fn(shard, idx):
    seen.extend(idx)
    return [f'{s}@{dist.get_rank()}' for s in shard]
out = dp_map(fn, items, lens)
if dist.get_rank() == 0:
    assert [o.split('@')[0] for o in out] == items, out
    assert {o.split('@')[1] for o in out} == {'0', '1'}
    print('DPMAP_OK', out)
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
'''


def test_dp_map_world_size_2_gloo(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29533', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert 'DPMAP_OK' in outs[0]


def test_weight_quantiser_matches_oracle_and_pack_roundtrip():
    """quantize_model: product quantiser (phi3_b200/quant.py) == oracle restatement; the packed 4-bit stream unpacks
    (with the kernel's nibble arithmetic, emulated here) to the same codes."""
    import torch
    import phi3_b200  # noqa
    from phi3_b200 import quant
    from oracle.phi3_oracle import quantize_q4g32, quantize_model_weights
    torch.manual_seed(0)
    w = (torch.randn(48, 256) * 0.02).to(torch.bfloat16)
    w[3, :64] = 0                                               # an all-zero group
    c, s, b = quant.quantize_w4g64(w)
    q, sc, bi = quantize_q4g32(w, prec='b200', group=64, dtype=torch.float64)
    assert (c.reshape(48, 4, 64) == q).all() and (s.float() == sc.squeeze(-1)).all() and (b.float() == bi.squeeze(-1)).all()
    deq = quant.dequantize_w4g64(c, s, b).float()
    ref = quantize_model_weights({'x.weight': w, 'n.weight': w[0]})
    assert (deq - ref['x.weight']).abs().max() <= 2 ** -8 * ref['x.weight'].abs().max()
    assert ref['n.weight'] is w[0] or (ref['n.weight'] == w[0]).all()         # 1-D (norm) weights are not quantised
    assert (ref['x.weight'] - w.float()).abs().max() <= 0.08 * w.float().abs().max()   # 4-bit: half a step of (max-min)/15
    packed, meta = quant.pack_w4g64(c, s, b)
    N, K = c.shape
    words = (packed.view(N, K // 128, 4, 4, 4).to(torch.int64) * (256 ** torch.arange(4))).sum(-1)
    rec = torch.zeros(N, K, dtype=torch.int64)
    for ch in range(K // 128):
        for t in range(4):
            for gi in range(2):
                for i in range(2):
                    wv = words[:, ch, t, gi * 2 + i]
                    for j in range(4):
                        k = 128 * ch + 64 * gi + 16 * t + 8 * i + 2 * j
                        rec[:, k], rec[:, k + 1] = (wv >> (4 * j)) & 0xF, (wv >> (4 * j + 16)) & 0xF
    assert (rec == c.to(torch.int64)).all()
    assert (meta[..., 0] == s).all() and (meta[..., 1] == b).all()


# ------------------------------------------------------------------ persistent decode-layer kernel: host logic
@pytest.mark.parametrize('kind,N,K,nh,nkv,hd', [(0, 384, 384, 0, 0, 0), (1, 2048, 384, 0, 0, 0), (0, 64, 8192, 0, 0, 0),
                                                (2, 1152, 384, 4, 4, 96), (3, 640, 384, 0, 0, 0)])
def test_mega_stream_order_reproduces_the_matmul(kind, N, K, nh, nkv, hd):
    """pack_reference (mirror of p3_mega_pack) consumed in the kernel's order (mega.emulate_phase: CTA -> K-block -> tile ->
    warp -> fragment, lane (g,t) pairing a0..a3 with x[n][kbase+16j+4t..]) gives x @ W^T for every tile kind."""
    import torch
    import phi3_b200  # noqa
    from phi3_b200 import mega
    g = torch.Generator().manual_seed(N + K)
    W = torch.randn(N, K, generator=g).to(torch.bfloat16)
    x = torch.randn(5, K, generator=g).to(torch.bfloat16)
    pk = mega.pack_reference(W, kind, nh, nkv, hd)
    assert pk.numel() == W.numel()
    (off, ids, mx), = mega.build_schedule([(kind, N, K)], 7)
    y = mega.emulate_phase(pk, kind, N, K, x, off, ids, nh, nkv, hd)
    assert torch.allclose(y, x.float() @ W.float().T, atol=1e-3)


def test_mega_schedule_balances_cumulative_bytes():
    import phi3_b200  # noqa
    from phi3_b200 import mega
    phases = [(mega.RESID, 3072, 3072), (mega.SWIGLU, 16384, 3072), (mega.RESID, 3072, 8192), (mega.QKV_ROPE, 9216, 3072)]
    sched = mega.build_schedule(phases, 148)
    load = [0] * 148
    for (kind, N, K), (off, ids, mx) in zip(phases, sched):
        MT, T, n_kblk, nkb_w = mega.dims(kind, N, K)
        assert sorted(ids.tolist()) == list(range(T)) and off[-1] == T          # every tile exactly once
        cost = MT * 16 * K * 2
        for c in range(148):
            load[c] += (int(off[c + 1]) - int(off[c])) * cost
        assert max(load) - min(load) <= cost                                    # within one tile at every phase boundary
        assert K <= mega.KBLOCK or mx <= mega.MAX_PART
    assert sum(load) == sum(N * K * 2 for _, N, K in phases)
    assert max(load) / (sum(load) / 148) < 1.06


def test_mega_args_layout_matches_header():
    import ctypes
    import phi3_b200  # noqa
    from phi3_b200 import mega
    assert ctypes.sizeof(mega.MegaPhase) == 104 and ctypes.sizeof(mega.MegaArgs) == 520    # static_assert in decode_mega.cu
    from phi3_b200 import _lib
    assert ctypes.sizeof(_lib.SkinnyArgs) == 272 and _lib.SkinnyArgs.past_dev.offset == 240 and _lib.SkinnyArgs.packed.offset == 264  # static_assert in gemm_skinny.cu
    assert ctypes.sizeof(_lib.GemmArgs) == 224


# ------------------------------------------------------------------ round 2 host logic: batching, checkpoint directories
def _fake_vproc():
    import torch
    import phi3_b200  # noqa
    from phi3_b200.processor import Phi3VProcessor, ByteTokenizer

    class FakeIP:
        num_crops = 4

        def __call__(self, images):
            n = len(images)
            return {'pixel_values': torch.zeros(n, 5, 3, 336, 336), 'image_sizes': [[672, 672]] * n, 'num_img_tokens': [757] * n}
    proc = Phi3VProcessor.__new__(Phi3VProcessor)
    proc.tokenizer, proc.img_processor = ByteTokenizer(), FakeIP()
    return proc


def test_batch_inputs_keep_per_prompt_batch1_semantics():
    """api._batch_inputs: every row is the batch-1 processor output, left-padded (ids 0 / pids 1 / mask 0, phi:236-245) with
    the image-token positions shifted by the row's pad (H11)."""
    import torch
    from phi3_b200 import api
    proc = _fake_vproc()
    img = torch.zeros(672, 672, 3, dtype=torch.uint8)
    texts = [api._apply_chat_template(t, im, False, True)[0] for t, im in (('a' * 100, [img]), ('hello', None), ('x' * 10, [img]))]
    b = api._batch_inputs(proc, texts, [[img], None, [img]])
    L = b['input_ids'].shape[1]
    assert api._prompt_lengths(proc, texts, [[img], None, [img]]) == b['mask'].sum(1).tolist()
    for r, (t, im) in enumerate(zip(texts, ([img], None, [img]))):
        one = proc(t, im) if im else proc(t)
        l = one['input_ids'].shape[1]
        assert b['input_ids'][r, L - l:].tolist() == one['input_ids'][0].tolist()
        assert b['input_ids'][r, :L - l].eq(0).all() and b['pids'][r, :L - l].eq(1).all() and b['mask'][r, :L - l].eq(0).all()
        assert b['pids'][r, L - l:].tolist() == list(range(l))
    pos = b['positions']
    assert pos.shape[0] == 2 * 757 and set(pos[:, 0].tolist()) == {0, 2}
    assert (b['input_ids'][pos[:, 0], pos[:, 1]] < 0).all() and (b['input_ids'] < 0).sum() == 2 * 757
    assert b['pixel_values'].shape[0] == 2 and b['image_sizes'].tolist() == [[672, 672], [672, 672]]


def test_checkpoint_directory_config_and_mlx_quantized_layout(tmp_path):
    """_read_config mirrors _get_cfg (pv:359-369); _read_safetensors handles shards, the HF patch-conv layout (pv:374), the
    `sanitized` flag (pv:276-289) and MLX's quantized weight/scales/biases triples (pv:291-305)."""
    import json
    import torch
    from safetensors.torch import save_file
    import phi3_b200  # noqa
    from phi3_b200 import api
    d = tmp_path / 'ckpt'
    d.mkdir()
    cfg = dict(architectures=['Phi3VForCausalLM'], hidden_size=64, num_attention_heads=2, rope_theta=12345.0,
               rope_scaling={'type': 'su', 'short_factor': [1.0] * 16, 'long_factor': [2.0] * 16})
    json.dump(cfg, open(d / 'config.json', 'w'))
    key = 'model.vision_embed_tokens.img_processor.vision_model.embeddings.patch_embedding.weight'
    conv = torch.randn(8, 3, 14, 14).to(torch.bfloat16)
    save_file({key: conv}, str(d / 'model-00001-of-00002.safetensors'))
    save_file({'lm_head.weight': torch.randn(4, 64).to(torch.bfloat16)}, str(d / 'model-00002-of-00002.safetensors'))
    c = api._read_config(str(d))
    assert c.rope_theta == 12345.0 and c.num_key_value_heads == 2 and c.use_quantized_cache is False
    w = api._read_safetensors(str(d))
    assert set(w) == {key, 'lm_head.weight'} and w[key].shape == (8, 14, 14, 3)
    assert torch.equal(w[key], conv.permute(0, 2, 3, 1))
    # sanitize(): reference layout on disk + sanitized flag; reading it back must NOT transpose again
    api.sanitize(str(d), str(tmp_path / 'san'))
    c2 = api._read_config(str(tmp_path / 'san'))
    assert c2.sanitized is True
    w2 = api._read_safetensors(str(tmp_path / 'san'), sanitized=True)
    assert torch.equal(w2[key], w[key])
    assert api._read_config(str(tmp_path)) is None
    # MLX quantized triple: element i of a row sits in bits [4*(i%8), +4) of uint32 word i//8
    q = torch.randint(0, 16, (4, 128))
    words = (q.reshape(4, 16, 8).to(torch.int64) << (4 * torch.arange(8))).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)       # safetensors has no uint32 in torch
    sc, bi = torch.rand(4, 2) + 0.1, torch.randn(4, 2)
    ref = (q.reshape(4, 2, 64).float() * sc[..., None] + bi[..., None]).reshape(4, 128)
    got = api.dequantize_mlx(words, sc, bi, 64, 4)
    assert torch.allclose(got.float(), ref, rtol=1e-2, atol=1e-2)


def test_logit_stopper_accepts_bool_like_the_reference():
    from phi3_b200 import api
    assert api.LogitStopper(100, True).early_stop is True            # pv:82: bool is an int -> enabled with threshold 1
    assert api.LogitStopper(100, False).early_stop is False
    assert api.LogitStopper(100, 5).early_stop == 5 and api.LogitStopper(4, 5).early_stop is False


def test_logit_stopper_decisions_equal_the_oracle_restatement():
    """api.LogitStopper (fed the two scalars the device computes) against oracle.drivers.LogitStopper (pv:79-104, fed full
    logits) on scripted sequences, including one that fires the early stop."""
    import torch
    from phi3_b200 import api
    from oracle import drivers
    V, EOS = 32064, 32007
    g = torch.Generator().manual_seed(0)
    n_fired = 0
    for trial in range(6):
        a, o = api.LogitStopper(64, 3), drivers.LogitStopper(64, 3)
        fired = None
        for step in range(40):
            logits = torch.randn(1, 1, V, generator=g)
            logits[0, 0, 5] += 8.0                                  # a clear best token
            logits[0, 0, EOS] += (step * 0.6 if trial % 2 == 0 else -3.0)   # EOS log-prob keeps rising in even trials
            lp = torch.log_softmax(logits[:, -1, :], -1)
            ra = a(lp.max().item(), lp[0, EOS].item())
            ro = o(logits)
            assert ra == ro, (trial, step)
            if ra:
                fired = step
                break
        n_fired += fired is not None
    assert n_fired >= 1                                             # the stop path was exercised


# ------------------------------------------------------------------ batching HTTP front end (srv:1-38)
def _post(port, payload, path='/v1/completions'):
    import json
    import urllib.request
    req = urllib.request.Request(f'http://127.0.0.1:{port}{path}', data=json.dumps(payload).encode(),
                                 headers={'Content-Type': 'application/json'})
    with urllib.request.urlopen(req, timeout=30) as r:
        return r.status, json.loads(r.read().decode())


def test_server_wire_contract_and_request_batching():
    """Same endpoint / JSON as the reference server; concurrent requests share batched generate calls and every client gets
    its own answers back in order; a request with another max_tokens never joins the batch; unknown paths are 404."""
    import threading
    import time
    import urllib.error
    import phi3_b200  # noqa
    from phi3_b200 import server
    seen = []

    def fake_generate(prompts, max_tokens):
        seen.append((list(prompts), max_tokens))
        time.sleep(0.15)                                     # a "decode" long enough for the other clients to queue up
        return [f'{p}|{max_tokens}' for p in prompts]
    httpd = server.serve(fake_generate, port=0, max_batch=8, window_ms=30, host='127.0.0.1')
    port = httpd.server_address[1]
    t = threading.Thread(target=httpd.serve_forever, daemon=True)
    t.start()
    try:
        st, body = _post(port, {'prompt': 'solo', 'max_tokens': 5})
        assert st == 200 and body == {'model': 'phi-3-vision', 'responses': ['solo|5']}     # str in -> list of one (srv:15-16,21-22)
        st, body = _post(port, {'prompt': ['a', 'b']})
        assert body['responses'] == ['a|512', 'b|512']                                       # default max_tokens (srv:14)
        n0 = len(seen)
        results = {}

        def client(i):
            mt = 7 if i == 5 else 9
            results[i] = _post(port, {'prompt': [f'c{i}x', f'c{i}y'] if i % 2 else f'c{i}', 'max_tokens': mt})[1]['responses']
        th = [threading.Thread(target=client, args=(i,)) for i in range(6)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        for i in range(6):
            mt = 7 if i == 5 else 9
            want = [f'c{i}x|{mt}', f'c{i}y|{mt}'] if i % 2 else [f'c{i}|{mt}']
            assert results[i] == want
        calls = seen[n0:]
        assert len(calls) < 6                                # coalesced
        assert all(len(p) <= 8 for p, _ in calls)            # max_batch respected
        assert all(len({mt}) == 1 for _, mt in calls) and any(mt == 7 and len(p) == 2 for p, mt in calls)
        with pytest.raises(urllib.error.HTTPError) as e:
            _post(port, {'prompt': 'x'}, path='/v1/other')
        assert e.value.code == 404
    finally:
        httpd.shutdown()
        httpd.batcher.close()


def test_server_failure_reaches_every_waiting_client_and_worker_survives():
    import threading
    import urllib.error
    import phi3_b200  # noqa
    from phi3_b200 import server

    def flaky(prompts, max_tokens):
        if any('boom' in p for p in prompts):
            raise RuntimeError('decode failed')
        return [p.upper() for p in prompts]
    httpd = server.serve(flaky, port=0, max_batch=4, window_ms=1, host='127.0.0.1')
    port = httpd.server_address[1]
    threading.Thread(target=httpd.serve_forever, daemon=True).start()
    try:
        with pytest.raises(urllib.error.HTTPError) as e:
            _post(port, {'prompt': 'boom'})
        assert e.value.code == 500
        assert _post(port, {'prompt': 'ok'})[1]['responses'] == ['OK']
        b = server.Batcher(lambda p, m: p[:-1], max_batch=4, window_ms=1)     # wrong number of answers is an error, not a mix-up
        with pytest.raises(RuntimeError):
            b.submit(['a', 'b'], 3)
        b.close()
    finally:
        httpd.shutdown()
        httpd.batcher.close()


def test_pack_rows16_layout_matches_the_header_formula():
    """p3_skinny_args.packed (include/phi3_b200.h): [N/16 tiles][K/64 chunks][row half][k half][lane = 4 * (row % 8) + quad][8]"""
    import torch
    import phi3_b200  # noqa
    from phi3_b200.model import pack_rows16
    N, K = 48, 192
    w = torch.arange(N * K, dtype=torch.float32).view(N, K)
    p = pack_rows16(w).view(-1)
    for n, k in [(0, 0), (7, 63), (8, 0), (15, 191), (16, 64), (37, 100), (47, 191)]:
        flat = (((n // 16) * (K // 64) + k // 64) * 4 + ((n % 16) // 8) * 2 + (k % 64) // 32) * 256 + (4 * (n % 8) + (k % 32) // 8) * 8 + k % 8
        assert p[flat] == w[n, k]
    assert torch.equal(p.sort().values, w.view(-1))                  # a permutation
    with pytest.raises(AssertionError):
        pack_rows16(torch.zeros(40, 192))
