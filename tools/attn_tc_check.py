"""Diagnostics for the tcgen05 flash-attention kernel: error breakdown vs a torch fp32 restatement and
vs the mma.sync kernel, then timings of both on the bench shapes.  python tools/attn_tc_check.py [--time]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phi3_b200 import _lib as L

dev = torch.device('cuda')
bf = lambda x: x.to(torch.bfloat16)
st = lambda: torch.cuda.current_stream().cuda_stream


def paged(kc, vc):
    B, H, S, D = kc.shape
    pps = (S + 63) // 64
    pool = torch.zeros(B * pps, 2, H, 64, D, device=dev, dtype=torch.bfloat16)
    bt = torch.arange(B * pps, dtype=torch.int32, device=dev).reshape(B, pps).flip(1).contiguous()
    for b in range(B):
        for p in range(pps):
            n = min(64, S - p * 64)
            pool[bt[b, p], 0, :, :n] = kc[b, :, p * 64:p * 64 + n]
            pool[bt[b, p], 1, :, :n] = vc[b, :, p * 64:p * 64 + n]
    return pool, bt


def ref(q, k, v, scale, causal, past, kv_start):
    B, H, Lq, D = q.shape
    S = k.shape[2]
    s = (q.float() * scale) @ k.float().transpose(-1, -2)
    qi = past + torch.arange(Lq, device=dev)[:, None]
    kj = torch.arange(S, device=dev)[None, :]
    allow = (kj <= qi) if causal else torch.ones(Lq, S, dtype=torch.bool, device=dev)
    allow = allow[None, None] & (kj[None, None] >= kv_start[:, None, None, None])
    s = s.masked_fill(~allow, float('-inf'))
    dead = (~allow).all(-1, keepdim=True)
    p = torch.softmax(s.masked_fill(dead, 0), -1).masked_fill(dead, 0)
    return p @ v.float()


def run(B, H, D, Lq, past, causal, kvs, tc, qkv, kc, vc):
    os.environ['P3_ATTN_TC'] = '1' if tc else '0'
    out = torch.full((B * Lq, H * D), float('nan'), device=dev, dtype=torch.bfloat16)
    pool, bt = paged(kc, vc) if past else (None, None)
    p = qkv.data_ptr()
    L.call('p3_attention_prefill', p, p + H * D * 2, p + 2 * H * D * 2, 3 * H * D, 3 * H * D, 3 * H * D, out.data_ptr(),
           H * D, B, Lq, H, H, D, D ** -0.5, int(causal), past, kvs.data_ptr(),
           None if pool is None else pool.data_ptr(), None if bt is None else bt.data_ptr(), 0 if bt is None else bt.stride(0), 1, st())
    torch.cuda.synchronize()
    return out


def check(cfg):
    B, H, D, Lq, past, causal, kv = cfg[:7]
    torch.manual_seed(3)
    qkv = bf(torch.randn(B * Lq, 3 * H * D, device=dev))
    kc, vc = bf(torch.randn(B, H, past, D, device=dev)), bf(torch.randn(B, H, past, D, device=dev))
    kvs = torch.tensor(kv[:B], dtype=torch.int32, device=dev)
    if len(cfg) > 7:                                    # growing score magnitude: exercises the lazy-rescale and redo paths
        kview = qkv.view(B, Lq, 3, H, D)[:, :, 1]
        ramp = torch.tensor([cfg[7] ** (i // 128) for i in range(Lq)], device=dev, dtype=torch.float32)
        kview.mul_(ramp[None, :, None, None].to(torch.bfloat16))
    x = qkv.view(B, Lq, 3, H, D).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], torch.cat([kc, x[1]], 2), torch.cat([vc, x[2]], 2)
    r = ref(q, k, v, D ** -0.5, causal, past, kvs.long()).transpose(1, 2).reshape(B, Lq, H, D)
    res = {}
    for tc in (1, 0):
        o = run(B, H, D, Lq, past, causal, kvs, tc, qkv, kc, vc).float().view(B, Lq, H, D)
        e = (o - r)
        den = r.abs().max().item()
        line = f'  tc={tc} max_rel={e.abs().max().item() / den:.3e} nan={int(torch.isnan(o).sum())}'
        if tc:
            line += f' | dims0-31 {e[..., :32].abs().max().item() / den:.2e} 32-63 {e[..., 32:64].abs().max().item() / den:.2e}'
            if D > 64:
                line += f' 64-95 {e[..., 64:].abs().max().item() / den:.2e}'
            blocks = [f'{torch.nan_to_num(e[:, i:i + 128], nan=9.9).abs().max().item() / den:.1e}' for i in range(0, Lq, 128)]
            line += ' | row blocks ' + ' '.join(blocks[:8])
        print(line, flush=True)
        res[tc] = e.abs().max().item() / den
    return res


def timeit(B, H, D, Lq, causal, tc, iters=10):
    os.environ['P3_ATTN_TC'] = '1' if tc else '0'
    qkv = bf(torch.randn(B * Lq, 3 * H * D, device=dev))
    out = torch.empty(B * Lq, H * D, device=dev, dtype=torch.bfloat16)
    kvs = torch.zeros(B, dtype=torch.int32, device=dev)
    p = qkv.data_ptr()
    def go():
        L.call('p3_attention_prefill', p, p + H * D * 2, p + 2 * H * D * 2, 3 * H * D, 3 * H * D, 3 * H * D, out.data_ptr(),
               H * D, B, Lq, H, H, D, D ** -0.5, int(causal), 0, kvs.data_ptr(), None, None, 0, 1, st())
    for _ in range(3):
        go()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        go()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    flops = 4.0 * B * H * Lq * Lq * D * (0.5 if causal else 1.0)
    return ms, flops / ms / 1e9


if __name__ == '__main__':
    cfgs = [(1, 1, 64, 128, 0, False, [0, 0, 0]), (1, 1, 96, 128, 0, False, [0, 0, 0]), (1, 1, 96, 128, 0, True, [0, 0, 0]),
            (1, 2, 96, 256, 0, True, [0, 0, 0]), (2, 4, 96, 200, 0, True, [0, 17, 70]), (1, 16, 64, 577, 0, False, [0, 0, 0]),
            (2, 3, 96, 300, 256, True, [0, 200, 0]), (1, 2, 96, 640, 128, True, [5, 0, 0]), (3, 2, 64, 577, 0, False, [0, 0, 0]),
            (1, 4, 96, 2048, 0, True, [0, 0, 0]), (2, 2, 96, 1100, 0, True, [0, 300, 0]),
            (1, 2, 96, 1024, 0, True, [0, 0, 0], 3.0), (1, 2, 96, 768, 0, False, [0, 0, 0], 40.0), (2, 2, 64, 700, 0, False, [0, 0, 0], 6.0)]
    bad = 0
    for c in cfgs:
        print('cfg (B,H,D,L,past,causal,kv_start)=', c, flush=True)
        r = check(c)
        bad += r[1] > 2e-2 or r[1] != r[1]
    print('FAILED' if bad else 'ALL OK', bad)
    if '--time' in sys.argv:
        for name, shp in (('prefill B8 H32 L2048 D96 causal', (8, 32, 96, 2048, True)), ('ViT 40 crops H16 L577 D64', (40, 16, 64, 577, False)),
                          ('prefill B1 H32 L8192 D96 causal', (1, 32, 96, 8192, True))):
            for tc in (1, 0):
                ms, tf = timeit(*shp, tc)
                print(f'{name}: tc={tc} {ms:.3f} ms {tf:.0f} TFLOP/s', flush=True)
