"""BASELINE configs[4] alone (the `constrain_cfg5` extra of bench.py): quantize_cache + constrain(use_beam, n_beam=4), 16 prompts of
250-450 tokens, constraint strings at their Phi-3 token counts. Seconds per call (best of 3); for same-box A/B of knobs:
    [P3_GEMM_BN64=0] [P3_ATTN_EARLY=0] [P3_XG=0] python tools/cfg5_bench.py"""
import os
import sys
import time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import configs, weights, api
from phi3_b200.model import Phi3B200
from phi3_b200.processor import ByteTokenizer, Phi3FProcessor


class Tok(ByteTokenizer):
    FIXED = {'\nThe': [29871, 13, 1576], ' The correct answer is': [29871, 450, 1959, 1234, 338]}

    def encode(self, text, add_special_tokens=True):
        return self.FIXED.get(text) or super().encode(text, add_special_tokens)


def main():
    dev = torch.device('cuda:0')
    cfg = configs.PHI35_MINI
    model = Phi3B200(cfg, weights.random_weights(cfg, seed=0, device=dev), device=dev)
    rs = np.random.RandomState(5)
    qs = [''.join(chr(c) for c in rs.randint(97, 123, int(n))) for n in rs.randint(250, 451, 16)]
    prompts = api._apply_chat_template(qs, None, False)[0]
    proc = Phi3FProcessor(Tok())
    cons = [(0, '\nThe'), (100, ' The correct answer is'), 'ABCDE']
    model.use_quantized_cache = True
    model.cfg.allow_beam_with_quantized_cache = True
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = api._constrain(model, proc, prompts, cons, mute=True, verbose=False, use_beam=True, n_beam=4)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    knobs = {k: os.environ[k] for k in ('P3_GEMM_BN64', 'P3_ATTN_EARLY', 'P3_XG', 'P3_SK_PACK', 'P3_SK_GRID') if k in os.environ}
    print(f'{min(ts):.4f} s per call (runs: {", ".join(f"{t:.3f}" for t in ts)})  rows={len(out)} {knobs}')


if __name__ == '__main__':
    main()
