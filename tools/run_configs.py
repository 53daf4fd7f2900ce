"""One timed pass of each BASELINE config that is not the bench workload (parity for them lives in tests/):
    python tools/run_configs.py [1] [2] [4] [5]   -> JSON lines
cfg1 Phi-3.5-mini 4 prompts x 128 tokens; cfg2 single-image VQA (num_crops=4) prefill + 128 tokens;
cfg4 128K-token prompt (long LongRoPE factors, chunked prefill, 51.5 GB bf16 KV) + 128 tokens;
cfg5 quantize_cache + constrained beam decoding, batch 16, beam 4, MedQA-shaped prompts."""
import json
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import configs, weights, api
from phi3_b200.model import Phi3B200
from phi3_b200.processor import ByteTokenizer, Phi3FProcessor, Phi3VImageProcessor, hd_geometry
from phi3_b200.api import _row_stats

dev = torch.device('cuda:0')


def sync_time(fn, n=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, out


def gen(model, ids, new, **kw):
    lg, cache = model(ids, max_tokens=new, logits_rows='last', **kw)
    tok = _row_stats(model, lg[:, -1, :])['argmax']
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    hist = model.greedy_decode(tok, cache, new - 1)
    torch.cuda.synchronize()
    return time.perf_counter() - t1, hist


def main():
    args = sys.argv[1:]
    which = [int(a) for a in args if a.isdigit()] or ([] if 'w4' in args else [1, 2, 4, 5])
    g = torch.Generator().manual_seed(0)
    if 'w4' in args:
        # quantize_model=True (4-bit g64 weights, pv:264,291-305) on config 1 and on the bench's decode shape
        cfg = configs.PHI35_MINI
        qm = Phi3B200(cfg, weights.random_weights(cfg, seed=0, device=dev), device=dev, quantize_model=True)
        ids = torch.randint(3, 32000, (4, 32), generator=g); ids[:, 0] = 1
        for _ in range(2):
            t_all, (t_dec, hist) = sync_time(lambda: gen(qm, ids, 128))
        print(json.dumps({'config': '1 + quantize_model', 'desc': 'Phi-3.5-mini 4-bit g64 weights, 4 prompts x 32 ctx x 128 new',
                          'total_ms': round(t_all * 1e3, 2), 'decode_tok_per_s': round(4 * 127 / t_dec, 1),
                          'decode_ms_per_token': round(t_dec / 127 * 1e3, 3)}), flush=True)
        ids = torch.randint(3, 32000, (8, 2048), generator=g); ids[:, 0] = 1
        for _ in range(2):
            t_all, (t_dec, hist) = sync_time(lambda: gen(qm, ids, 256))
        print(json.dumps({'config': '3 (text-only shape) + quantize_model', 'desc': '8 x 2048 ctx x 256 new, 4-bit g64 weights, bf16 KV',
                          'total_ms': round(t_all * 1e3, 2), 'decode_tok_per_s': round(8 * 255 / t_dec, 1),
                          'decode_ms_per_token': round(t_dec / 255 * 1e3, 3)}), flush=True)
        ids1 = torch.randint(3, 32000, (1, 32), generator=g); ids1[:, 0] = 1
        for _ in range(2):
            t_all, (t_dec, hist) = sync_time(lambda: gen(qm, ids1, 128))
        print(json.dumps({'config': 'single stream + quantize_model', 'desc': '1 prompt x 32 ctx x 128 new (the reference table\'s case)',
                          'decode_tok_per_s': round(127 / t_dec, 1), 'decode_ms_per_token': round(t_dec / 127 * 1e3, 3)}), flush=True)
        del qm
        torch.cuda.empty_cache()
    if 1 in which or 4 in which or 5 in which:
        cfg = configs.PHI35_MINI
        mini = Phi3B200(cfg, weights.random_weights(cfg, seed=0, device=dev), device=dev)
    if 1 in which:
        ids = torch.randint(3, 32000, (4, 32), generator=g); ids[:, 0] = 1
        for _ in range(2):
            t_all, (t_dec, hist) = sync_time(lambda: gen(mini, ids, 128))
        print(json.dumps({'config': 1, 'desc': 'Phi-3.5-mini greedy, 4 prompts x 32 ctx x 128 new', 'total_ms': round(t_all * 1e3, 2),
                          'decode_tok_per_s': round(4 * 127 / t_dec, 1), 'decode_ms_per_token': round(t_dec / 127 * 1e3, 3)}))
    if 4 in which:
        L = 130944
        ids = torch.randint(3, 32000, (1, L), generator=g); ids[:, 0] = 1
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lg, cache = mini(ids, max_tokens=128, logits_rows='last')
        tok = _row_stats(mini, lg[:, -1, :])['argmax']
        torch.cuda.synchronize(); t_pre = time.perf_counter() - t0
        t1 = time.perf_counter()
        hist = mini.greedy_decode(tok, cache, 127)
        torch.cuda.synchronize(); t_dec = time.perf_counter() - t1
        kv_gb = cache.pool.numel() * 2 / 1e9
        print(json.dumps({'config': 4, 'desc': '128K single sequence, long SuRoPE factors, chunked prefill 8192', 'context': L,
                          'prefill_s': round(t_pre, 2), 'prefill_tok_per_s': round(L / t_pre, 0), 'kv_pool_gb': round(kv_gb, 1),
                          'decode_tok_per_s': round(127 / t_dec, 1), 'decode_ms_per_token': round(t_dec / 127 * 1e3, 2),
                          'decode_kv_gbs': round(L * 393216 / (t_dec / 127) / 1e9, 0)}))
        del cache
        mini._slabs.clear()
    if 5 in which:
        cfgq = configs.with_overrides(configs.PHI35_MINI, use_quantized_cache=True, allow_beam_with_quantized_cache=True)
        mini.cfg, mini.use_quantized_cache = cfgq, True
        tok = ByteTokenizer()
        proc = Phi3FProcessor(tok)
        import random
        random.seed(0)
        prompts = [''.join(random.choice('abcdefghij klmnop') for _ in range(random.randint(250, 450))) for _ in range(16)]
        cons = [(0, '\nThe'), (20, ' The correct answer is'), 'ABCDE']
        for nb in (3, 4):
            t, out = sync_time(lambda: api._constrain(mini, proc, prompts, cons, mute=True, verbose=False, use_beam=True, n_beam=nb))
            print(json.dumps({'config': 5, 'desc': f'quantize_cache + constrain use_beam n_beam={nb}, batch 16, prompts 250-450 tokens, '
                              'constraints [(0,..),(20,..),ABCDE]', 'total_s': round(t, 2)}))
        mini.cfg, mini.use_quantized_cache = configs.PHI35_MINI, False
    if 2 in which:
        cfg = configs.PHI35_VISION
        vis = Phi3B200(cfg, weights.random_weights(cfg, seed=0, device=dev), device=dev)
        ip = Phi3VImageProcessor(num_crops=4, device=dev)
        geo = hd_geometry(672, 672, 4)
        n = geo['num_img_tokens']
        img = torch.randint(0, 256, (672, 672, 3), generator=g, dtype=torch.uint8).to(dev)
        ids = torch.randint(3, 32000, (1, n + 24), generator=g); ids[:, 0] = 1; ids[:, 6:6 + n] = -1
        pos = torch.nonzero(ids < 0)
        sizes = torch.tensor([[geo['H'], geo['W']]])

        def vqa():
            pv = ip([img])['pixel_values']
            return gen(vis, ids, 128, pixel_values=pv, image_sizes=sizes, positions=pos)
        for _ in range(3):
            t_all, (t_dec, hist) = sync_time(vqa)
        print(json.dumps({'config': 2, 'desc': 'Phi-3.5-vision single 672x672 image, num_crops=4 (757 image tokens), 781-token prompt, 128 new',
                          'prefill_ms_incl_hd_transform': round((t_all - t_dec) * 1e3, 2), 'decode_tok_per_s': round(127 / t_dec, 1),
                          'decode_ms_per_token': round(t_dec / 127 * 1e3, 3)}))


if __name__ == '__main__':
    main()
