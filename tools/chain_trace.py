"""Timeline of the decode launch chain from in-kernel globaltimer stamps (p3_trace_set): for every launch of a few layers of one
eager decode step (PDL chain, GPU parked first so the host is ahead) and of one CUDA-graph replay: when its first / last CTA
started, when the dependency wait released, when the first operands had landed, when the main loops ended and when the last CTA
left — i.e. where HBM idles at the kernel boundaries.
    python tools/chain_trace.py [--B 8] [--ctx 2048] [--layers 3] [--qm]"""
import argparse
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import configs, weights, _lib
from phi3_b200.model import Phi3B200

KIND = {900: 'attn'}


def kind_name(k, grid):
    if k == 900:
        return 'attn'
    epi, mt = k // 100, (k // 10) % 10
    return {0: 'none', 3: 'resid', 4: 'swiglu', 5: 'f32', 7: 'qkv+rope'}.get(epi, str(epi)) + f'/mt{mt}'


def report(tr, n, first, count, title):
    print(f'--- {title}: launches {first}..{first + count - 1}')
    print(f'{"#":>4} {"kernel":14} {"CTAs":>5} | {"first start":>11} {"last start":>10} | {"wait rel. min":>13} {"max":>7} | '
          f'{"1st data med":>12} | {"loop end min":>12} {"med":>7} {"max":>7} | {"exit max":>8} | {"prev exit->rel":>14} {"rel->1st data":>13}')
    t0 = None
    prev_exit = None
    for i in range(first, min(first + count, n)):
        g = int(tr[i, 0, 7])
        if g == 0:
            continue
        g = min(g, 1024)
        a = tr[i, :g].astype(np.int64)
        if t0 is None:
            t0 = a[:, 0].min()
        us = lambda x: (x - t0) / 1e3
        st, rel, dat, le, ex = a[:, 0], a[:, 1], a[:, 2], a[:, 3], a[:, 4]
        gap = '' if prev_exit is None else f'{(rel.min() - prev_exit) / 1e3:14.2f}'
        print(f'{i:4d} {kind_name(int(a[0, 6]), g):14} {g:5d} | {us(st.min()):11.2f} {us(st.max()):10.2f} | {us(rel.min()):13.2f} {us(rel.max()):7.2f} | '
              f'{us(np.median(dat)):12.2f} | {us(le.min()):12.2f} {us(np.median(le)):7.2f} {us(le.max()):7.2f} | {us(ex.max()):8.2f} | {gap:>14} '
              f'{(np.median(dat) - rel.min()) / 1e3:13.2f}')
        prev_exit = ex.max()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=8)
    ap.add_argument('--ctx', type=int, default=2048)
    ap.add_argument('--layers', type=int, default=3)
    ap.add_argument('--qm', action='store_true')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    cfg = configs.PHI35_MINI
    m = Phi3B200(cfg, weights.random_weights(cfg, seed=0, device=dev), device=dev, quantize_model=a.qm)
    ids = torch.randint(3, 32000, (a.B, a.ctx))
    ids[:, 0] = 1
    lg, c = m(ids, max_tokens=80, logits_rows='last')
    n_slots = 400
    buf = torch.zeros((n_slots, 1024, 8), dtype=torch.int64, device=dev)
    L = _lib.lib()
    # ---- graph replay: the trace pointer of every launch is frozen at capture time; each replay overwrites the same slots
    L.p3_trace_set(buf.data_ptr(), n_slots)
    ses = m.decode_session(lg[:, -1].argmax(-1).to(torch.int32), c, 60)
    for _ in range(6):
        ses.step()
    torch.cuda.synchronize()
    n = L.p3_trace_count()
    L.p3_trace_set(None, 0)
    tr = buf.cpu().numpy()
    kinds = [int(tr[i, 0, 6]) for i in range(n)]
    qkv_at = [i for i, k in enumerate(kinds) if k // 100 == 7]
    print(f'{n} traced launches, {len(qkv_at)} qkv launches (decode attention is traced only in a -DP3_TRACE_ATTN build: its window is '
          f'qkv exit -> o_proj release)')
    last_step = qkv_at[-32:]
    first = last_step[0]
    lo = last_step[8]
    hi = last_step[8 + a.layers]
    report(tr, n, lo, hi - lo, f'last traced step (captured graph replay), layers 8..{8 + a.layers - 1}')
    report(tr, n, n - 6, 6, 'end of the step: last layer, lm_head')
    gl = (tr[first:n, 0, 7] > 0)
    exits = np.array([tr[i, :min(int(tr[i, 0, 7]), 1024), 4].max() for i in range(first, n) if tr[i, 0, 7] > 0])
    starts = np.array([tr[i, :min(int(tr[i, 0, 7]), 1024), 0].min() for i in range(first, n) if tr[i, 0, 7] > 0])
    print(f'step span (first start -> last exit): {(exits.max() - starts.min()) / 1e3:.1f} us over {gl.sum()} launches')


if __name__ == '__main__':
    main()
