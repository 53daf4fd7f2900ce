"""Decode-step time of the graph-replayed loop at the bench shape (8 rows x 2048 context), text-only Phi-3.5-mini weights:
    [P3_PF_CHAIN=0] [P3_PF_CAP_MB=..] [P3_SK_DEPTH=..] [P3_MEGA=1] python tools/decode_bench.py [--B 8] [--ctx 2048]
Prints ms per decode step (mean over 200 replays). For A/B of decode-path knobs inside ONE gpurun call."""
import argparse
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import configs, weights
from phi3_b200.model import Phi3B200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=8)
    ap.add_argument('--ctx', type=int, default=2048)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--qm', action='store_true', help='quantize_model (4-bit g64 weights)')
    ap.add_argument('--qc', action='store_true', help='quantize_cache (4-bit g32 KV)')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    cfg = configs.with_overrides(configs.PHI35_MINI, use_quantized_cache=a.qc)
    m = Phi3B200(cfg, weights.random_weights(cfg, seed=0, device=dev), device=dev, quantize_model=a.qm)
    ids = torch.randint(3, 32000, (a.B, a.ctx))
    ids[:, 0] = 1
    lg, c = m(ids, max_tokens=a.steps + 40, logits_rows='last')
    ses = m.decode_session(lg[:, -1].argmax(-1).to(torch.int32), c, a.steps + 30)
    for _ in range(20):
        ses.step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ses.step()
    e1.record()
    torch.cuda.synchronize()
    knobs = {k: os.environ[k] for k in ('P3_PF_CHAIN', 'P3_PF_CAP_MB', 'P3_SK_DEPTH', 'P3_MEGA', 'P3_OPF', 'P3_PRENORM', 'P3_PDL', 'P3_SKIP') if k in os.environ}
    print(f'{e0.elapsed_time(e1) / a.steps:.4f} ms/step  B={a.B} ctx={a.ctx} qm={a.qm} qc={a.qc} {knobs}')


if __name__ == '__main__':
    main()
