"""Micro-benchmark of the persistent decode-layer kernel alone (run on the GPU box):
    python tools/mega_bench.py [--B 8] [--reps 20]
31 'mid' launches (o_proj, gate_up, down, next qkv) back to back over the 31 layers' distinct weights (7 GB >> L2), one
CUDA-event pair around the lot; also the first ([qkv]) and last (... lm_head) launches. Prints JSON."""
import argparse
import json
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['P3_MEGA'] = '1'
import phi3_b200  # noqa
from phi3_b200 import configs, weights
from phi3_b200.model import Phi3B200, DecodeSession


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=8)
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--once', action='store_true', help='a single pass (for ncu)')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    cfg = configs.PHI35_MINI
    w = weights.random_weights(cfg, seed=0, device=dev)
    m = Phi3B200(cfg, w, device=dev)
    del w
    ids = torch.randint(3, 32000, (a.B, 64))
    lg, c = m(ids, max_tokens=8, logits_rows='last')
    ses = DecodeSession(m, lg[:, -1].argmax(-1), c, 4, use_graph=False)
    ms = ses.mega_ses
    st = torch.cuda.current_stream().cuda_stream
    nl = len(m.layers)
    sh = ms.mega.shapes
    per_layer = sum(N * K for k, (_, N, K) in sh.items() if k != 'lm') * 2

    def mid():
        for li in range(1, nl):
            ms.launch(li, st)
    mid()
    torch.cuda.synchronize()
    if a.once:
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        mid()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (a.reps * (nl - 1))
    out = {'B': a.B, 'mid_launch_us': round(us, 2), 'mid_bytes': per_layer, 'mid_GBps': round(per_layer / us / 1e3, 1)}
    e0.record()
    for _ in range(a.reps):
        ms.launch(nl, st)
        ms.launch(0, st)
    e1.record()
    torch.cuda.synchronize()
    us2 = 1e3 * e0.elapsed_time(e1) / a.reps
    b2 = per_layer + sh['lm'][1] * sh['lm'][2] * 2
    out.update(first_plus_last_us=round(us2, 2), first_plus_last_GBps=round(b2 / us2 / 1e3, 1))
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(peaks):
        pk = json.load(open(peaks))['hbm_gbs']
        out['frac_of_measured_hbm_peak'] = round(out['mid_GBps'] / pk, 4)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
