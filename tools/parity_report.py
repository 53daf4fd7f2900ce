"""Parity report at BASELINE config-1 shapes (run on the GPU box):
    python tools/parity_report.py [--layers N] > gpurun_out/parity.json
Phi-3.5-mini (random-init, seed 0), 4 prompts x 32 tokens; the CUDA path vs the CPU oracle in its
three precision modes. Reports logits error (max-abs/max-abs and rms/rms) and teacher-forced
greedy agreement over all positions, plus step-wise decode agreement through the skinny kernels."""
import argparse
import json
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import configs, weights
from phi3_b200.model import Phi3B200
from oracle.phi3_oracle import Phi3Oracle


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--layers', type=int, default=32)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--gen', type=int, default=96)
    ap.add_argument('--init', default='peaked', choices=['peaked', 'gpt2'],
                    help="'peaked' = the parity checkpoint (lm_head tied to the embedding); 'gpt2' = flat logits, measures the bf16 noise floor")
    a = ap.parse_args()
    cfg = configs.with_overrides(configs.PHI35_MINI, num_hidden_layers=a.layers)
    w = weights.random_weights(cfg, seed=0, init=a.init)
    m = Phi3B200(cfg, w)
    g = torch.Generator().manual_seed(11)
    ids = torch.randint(3, 32000, (4, 32), generator=g)
    ids[:, 0] = 1
    rep = {'config': f'Phi-3.5-mini {a.layers} layers, 4x32 prompt, random-init seed 0, init={a.init}', 'modes': {}}
    # greedy continuation from the CUDA path (device-resident loop), then teacher-force everything
    lg, cg = m(ids, max_tokens=a.gen, logits_rows='last')
    first = lg[:, -1].argmax(-1)
    hist = m.greedy_decode(first, cg, a.gen - 1).cpu().long()
    full = torch.cat([ids, hist[:, :-1]], 1)                      # [4, 32+gen-1]
    lgf, _ = m(full, max_tokens=0)
    lgf = lgf.cpu()
    gpu_arg = lgf.argmax(-1)
    rep['graph_decode_vs_gpu_prefill_agreement'] = (gpu_arg[:, 31:] == hist).float().mean().item()
    keep = {}
    for prec in ('b200', 'ref', 'fp32'):
        t0 = time.time()
        o = Phi3Oracle(cfg, w, prec=prec)
        lo, _ = o(full, max_tokens=0)
        keep[prec] = lo
        d = (lgf - lo)
        top2 = lo.topk(2, -1).values
        margin = (top2[..., 0] - top2[..., 1])
        agree = (gpu_arg == lo.argmax(-1))
        rep['modes'][prec] = {
            'max_abs_over_max_abs': (d.abs().max() / lo.abs().max()).item(),
            'rms_over_rms': (d.pow(2).mean().sqrt() / lo.pow(2).mean().sqrt()).item(),
            'teacher_forced_agreement': agree.float().mean().item(),
            'positions': int(agree.numel()),
            'median_top1_top2_margin': margin.median().item(),
            'agreement_where_margin_gt_0.1': agree[margin > 0.1].float().mean().item(),
            'positions_margin_gt_0.1': int((margin > 0.1).sum()),
            'margin_at_disagreements': margin[~agree].tolist()[:20],
            'oracle_seconds': time.time() - t0,
        }
    # the oracle against itself: how far the reference's own bf16 dtype flow is from ideal fp32 arithmetic
    for a_, b_ in (('ref', 'fp32'), ('b200', 'ref'), ('b200', 'fp32')):
        d = keep[a_] - keep[b_]
        rep['modes'][f'oracle_{a_}_vs_oracle_{b_}'] = {
            'rms_over_rms': (d.pow(2).mean().sqrt() / keep[b_].pow(2).mean().sqrt()).item(),
            'max_abs_over_max_abs': (d.abs().max() / keep[b_].abs().max()).item(),
            'greedy_agreement': (keep[a_].argmax(-1) == keep[b_].argmax(-1)).float().mean().item()}
    # step-wise decode through the skinny kernels, teacher-forced on the reference-flow oracle's tokens
    o = Phi3Oracle(cfg, w, prec='ref')
    lo, co = o(ids, max_tokens=a.steps + 1)
    lg, cg = m(ids, max_tokens=a.steps + 1)
    tok = lo[:, -1].argmax(-1)
    hit = tot = 0
    worst = 0.0
    for _ in range(a.steps):
        lo, co = o(tok[:, None], cache=co)
        lg, cg = m(tok[:, None], cache=cg)
        worst = max(worst, ((lg.cpu() - lo).abs().max() / lo.abs().max()).item())
        hit += (lg[:, -1].argmax(-1).cpu() == lo[:, -1].argmax(-1)).sum().item()
        tot += 4
        tok = lo[:, -1].argmax(-1)
    rep['stepwise_decode'] = {'steps': a.steps, 'agreement': hit / tot, 'worst_max_abs_over_max_abs': worst}
    print(json.dumps(rep, indent=1))


if __name__ == '__main__':
    main()
