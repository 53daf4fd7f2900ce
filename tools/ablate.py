"""Marginal cost of each decode-path kernel inside the PDL-chained CUDA graph: rerun the decode loop with
one kernel type skipped (P3_SKIP, timing only — outputs are garbage) and report the step-time delta."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, time, torch
sys.path.insert(0, %r)
import phi3_b200
from phi3_b200 import configs, weights
from phi3_b200.model import Phi3B200
dev = torch.device('cuda:0')
cfg = configs.PHI35_MINI
m = Phi3B200(cfg, weights.random_weights(cfg, seed=0, device=dev), device=dev)
ids = torch.randint(3, 32000, (8, 2048)); ids[:, 0] = 1
skip = m._skip; m._skip = set()
lg, c = m(ids, max_tokens=256, logits_rows='last')
m._skip = skip
tok = lg[:, -1].argmax(-1).to(torch.int32)
ses = m.decode_session(tok, c, 255)
for _ in range(20): ses.step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): ses.step()
torch.cuda.synchronize(); print((time.perf_counter() - t0) / 200 * 1e3)
''' % ROOT
res = {}
for skip in ['', 'qkv', 'attn', 'o', 'gu', 'down', 'o,down', 'qkv,attn,o,gu,down']:
    out = subprocess.run([sys.executable, '-c', CODE], env=dict(os.environ, P3_SKIP=skip), capture_output=True, text=True)
    try:
        res[skip or 'none'] = float(out.stdout.strip().split('\n')[-1])
    except Exception:
        res[skip or 'none'] = out.stderr[-300:]
base = res['none']
print(json.dumps({k: (round(v, 3), round(base - v, 3)) if isinstance(v, float) else v for k, v in res.items()}, indent=1))
